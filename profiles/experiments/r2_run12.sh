cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp12.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp12_tests.log 2>&1; echo "tests rc=$?" >> $L
timeout 200 python profiles/experiments/r2_lag_bench.py 2>&1 | grep -E "select|Error|error" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_exp12_bench.json 2> gpurun_out/r2_exp12_bench.err; echo "bench rc=$?" >> $L
tail -4 gpurun_out/r2_exp12_tests.log; tail -n 3 gpurun_out/r2_exp12_bench.err
cat $L
