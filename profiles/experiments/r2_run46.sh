cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp46.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp46_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp46_tests.log >> $L
timeout 600 python profiles/experiments/r2_cliffs.py 2>&1 | grep "localindex\]" >> $L
timeout 300 python profiles/experiments/r2_optable.py 'index_gather' >> $L 2>&1
cat $L
