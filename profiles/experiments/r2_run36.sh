cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp36.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "long_event or golden or selection_row or differential or api_cases" > gpurun_out/r2_exp36_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp36_tests.log >> $L
timeout 300 python profiles/experiments/r2_optable.py 'xxxx' --dist >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 60000000 10 2>&1 | grep "mean>7" >> $L
cat $L
