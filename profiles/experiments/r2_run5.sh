cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp5.log
: > $L
run() { name=$1; t=$2; shift 2; timeout $t "$@" > gpurun_out/r2_exp5_$name.log 2>&1; echo "$name rc=$?" >> $L; }
run tests 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_api_cases.py tests/test_gpu_api.py -q -x
JG_SCAN_MODE=0 timeout 200 python profiles/experiments/r2_lag_bench.py 2>&1 | grep -E "counts2offsets|Error|error" >> $L
for w in 512 1024; do for tk in 1 0; do
  echo "window=$w ticket=$tk" >> $L
  JG_SCAN_MODE=1 JG_SCAN_WINDOW=$w JG_TICKET=$tk timeout 200 python profiles/experiments/r2_lag_bench.py 2>&1 | grep -E "counts2offsets|Error|error" >> $L
done; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_exp5_bench.json 2> gpurun_out/r2_exp5_bench.err; echo "bench rc=$?" >> $L
tail -30 gpurun_out/r2_exp5_tests.log
tail -5 gpurun_out/r2_exp5_bench.err
cat $L
