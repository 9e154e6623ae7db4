set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_exp1.log 2>&1
# parity in the new modes (defaults: scan lag 1, select lag 1)
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_random.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r2_exp1_pytest_mode1.log 2>&1; echo "pytest mode1 rc=$?" >> gpurun_out/r2_exp1.log
JG_SCAN_MODE=2 JG_SELECT_MODE=2 timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_random.py -m gpu -x -q > gpurun_out/r2_exp1_pytest_mode2.log 2>&1; echo "pytest mode2 rc=$?" >> gpurun_out/r2_exp1.log
for m in 0 1 2; do
  JG_SCAN_MODE=$m JG_SELECT_MODE=$m timeout 300 python profiles/experiments/r2_lag_bench.py >> gpurun_out/r2_exp1.log 2>&1
done
for c in 2 4 6; do
  JG_SCAN_MODE=1 JG_SELECT_MODE=1 JG_SCAN_CTAS=$c JG_SELECT_CTAS=$c timeout 300 python profiles/experiments/r2_lag_bench.py >> gpurun_out/r2_exp1.log 2>&1
done
JG_SCAN_MODE=1 JG_SELECT_MODE=1 JG_SCAN_CTAS=8 JG_SELECT_CTAS=5 timeout 300 python profiles/experiments/r2_lag_bench.py >> gpurun_out/r2_exp1.log 2>&1
tail -5 gpurun_out/r2_exp1_pytest_mode1.log gpurun_out/r2_exp1_pytest_mode2.log
grep -v "^+" gpurun_out/r2_exp1.log
