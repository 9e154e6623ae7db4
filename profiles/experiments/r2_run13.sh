cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp13.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp13_tests.log 2>&1; echo "tests rc=$?" >> $L
timeout 300 python profiles/experiments/r2_pyprofile_sweep.py 2>&1 | head -3 >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_exp13_bench.json 2> gpurun_out/r2_exp13_bench.err; echo "bench rc=$?" >> $L
tail -4 gpurun_out/r2_exp13_tests.log; tail -n 3 gpurun_out/r2_exp13_bench.err
cat $L
