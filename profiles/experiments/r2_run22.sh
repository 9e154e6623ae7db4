cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp22.log
: > $L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_exp22_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -5 gpurun_out/r2_exp22_tests.log >> $L
cat $L
