cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp6.log
: > $L
run() { name=$1; t=$2; shift 2; timeout $t "$@" > gpurun_out/r2_exp6_$name.log 2>&1; echo "$name rc=$?" >> $L; }
run tests 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_api_cases.py tests/test_gpu_api.py -q -x
for w in 256 384 512 768; do
  echo "window=$w ticket=0" >> $L
  JG_SCAN_MODE=1 JG_SCAN_WINDOW=$w timeout 200 python profiles/experiments/r2_lag_bench.py 2>&1 | grep -E "counts2offsets|Error|error" >> $L
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_exp6_bench.json 2> gpurun_out/r2_exp6_bench.err; echo "bench rc=$?" >> $L
tail -30 gpurun_out/r2_exp6_tests.log
tail -5 gpurun_out/r2_exp6_bench.err
cat $L
