cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp47.log
: > $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> $L
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $L
tail -2 gpurun_out/r2_smoke.log >> $L
python bench.py --steps 2 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_plain_sweep1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sweep.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_ncu_sweep.log 2>&1; echo "ncu launches rc=$?" >> $L
cat $L
