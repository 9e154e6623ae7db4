cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp37.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp37_tests.log 2>&1; echo "tests rc=$?" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "reference rc=$?" >> $L
timeout 600 python bench.py --total-events 1e9 --steps 5 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_bench_1B_1gpu.json 2> gpurun_out/r2_bench_1B_1gpu.err; echo "1B rc=$?" >> $L
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $L
# DRAM traffic of the sweep's own kernels at 125 M events (same command plain first; no source import: small report)
python bench.py --steps 1 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_plain_sweep1.log 2>&1 && \
ncu --set full --clock-control none -k regex:'scan_win_kernel|segreduce_kernel|segarg_direct_kernel|mask_count_kernel|mask_fill_kernel' -s 6 -c 6 -f -o /tmp/r2_sweep python bench.py --steps 1 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_ncu_sweep_full.log 2>&1; echo "ncu sweep rc=$?" >> $L
ncu -i /tmp/r2_sweep.ncu-rep --page raw --csv > gpurun_out/r2_sweep_raw.csv 2>> $L
ncu -i /tmp/r2_sweep.ncu-rep --page source --csv > gpurun_out/r2_sweep_source.csv 2>> $L
# launch list of the sweep (the same command has just exited 0 without ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sweep.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_ncu_sweep.log 2>&1; echo "ncu launches rc=$?" >> $L
tail -4 gpurun_out/r2_exp37_tests.log; tail -3 gpurun_out/r2_bench_final.err gpurun_out/r2_bench_1B_1gpu.err gpurun_out/r2_smoke.log
head -c 600 gpurun_out/r2_bench_1B_1gpu.json; echo; head -c 1200 gpurun_out/r2_bench_final_reference.json; echo
cat $L
