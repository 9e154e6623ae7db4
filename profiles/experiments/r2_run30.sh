cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp30.log
: > $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 125000000 5 >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 60000000 10 >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 200000000 2 >> $L 2>&1; echo "rc=$?" >> $L
cat $L
