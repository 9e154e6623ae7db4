cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp2.log
: > $L
run() { # name timeout cmd...
  name=$1; t=$2; shift 2
  timeout $t "$@" > gpurun_out/r2_exp2_$name.log 2>&1; echo "$name rc=$?" >> $L
}
run big 300 python -m pytest tests/test_gpu_random.py -k "more_than_2_to_31 or one_event_longer" -x -q
run long 420 python -m pytest tests/test_gpu_random.py -k "long_event_variants or c1_full or c4_slice" -q
run api 300 python -m pytest tests/test_gpu_api_cases.py tests/test_gpu_api.py -q
for m in 0 1 2; do
  JG_SCAN_MODE=$m JG_SELECT_MODE=$m timeout 200 python profiles/experiments/r2_lag_bench.py >> $L 2>&1
done
for c in 3 5; do
  JG_SCAN_MODE=2 JG_SELECT_MODE=2 JG_SCAN_CTAS=$c JG_SELECT_CTAS=$c timeout 200 python profiles/experiments/r2_lag_bench.py >> $L 2>&1
done
for f in big long api; do echo "== $f"; tail -25 gpurun_out/r2_exp2_$f.log; done
cat $L
