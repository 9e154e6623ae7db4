cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp23.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "comb or pairs or cross or choose or distinct or lazy or api_cases or golden or physics" > gpurun_out/r2_exp23_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp23_tests.log >> $L
timeout 600 python profiles/experiments/r2_optable.py 'pairs|distincts|cross|choose' >> $L 2>&1
cat $L
