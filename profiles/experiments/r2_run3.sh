cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp3.log
: > $L
run() { name=$1; t=$2; shift 2; timeout $t "$@" > gpurun_out/r2_exp3_$name.log 2>&1; echo "$name rc=$?" >> $L; }
run tests 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_api_cases.py tests/test_gpu_api.py -q -x
for m in 0 1; do
  JG_SCAN_MODE=$m JG_SELECT_MODE=$m timeout 200 python profiles/experiments/r2_lag_bench.py >> $L 2>&1
done
for w in 512 4096; do
  JG_SCAN_MODE=1 JG_SELECT_MODE=1 JG_SCAN_WINDOW=$w JG_SELECT_WINDOW=$w timeout 200 python profiles/experiments/r2_lag_bench.py >> $L 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_exp3_bench.json 2> gpurun_out/r2_exp3_bench.err; echo "bench rc=$?" >> $L
tail -30 gpurun_out/r2_exp3_tests.log
tail -5 gpurun_out/r2_exp3_bench.err
cat $L
