cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp11.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp11_tests.log 2>&1; echo "tests rc=$?" >> $L
JG_SCAN_MODE=1 timeout 200 python profiles/experiments/r2_lag_bench.py 2>&1 | grep -E "counts2offsets|comb_offsets|Error|error" >> $L
timeout 300 python profiles/experiments/r2_pyprofile_sweep.py > gpurun_out/r2_pyprofile_sweep.txt 2>&1; echo "pyprofile rc=$?" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_exp11_bench.json 2> gpurun_out/r2_exp11_bench.err; echo "bench rc=$?" >> $L
tail -4 gpurun_out/r2_exp11_tests.log; head -3 gpurun_out/r2_pyprofile_sweep.txt; tail -n 3 gpurun_out/r2_exp11_bench.err
cat $L
