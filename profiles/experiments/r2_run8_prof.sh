cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp8.log
: > $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_exp8_bench.json 2> gpurun_out/r2_exp8_bench.err; echo "bench rc=$?" >> $L
# launch list of the sweep (same command plain first)
python bench.py --steps 2 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_plain_sweep.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sweep.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-tables > gpurun_out/r2_ncu_sweep.log 2>&1; echo "ncu launches rc=$?" >> $L
# full capture of the hot kernels on a 20 M-event shard: raw and source pages exported here, the report itself stays on the box
python profiles/prof_kernels.py > gpurun_out/r2_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_win_kernel|scan_tma_kernel|segreduce_kernel|segarg_direct_kernel|mask_count_kernel|mask_fill_kernel|index_gather_kernel|validate_offsets_kernel|segreduce_elem_kernel|segarg_elem_kernel|select_elem_fill_kernel|select_elem_count_kernel|expand_elem_kernel|expand_kernel' -s 30 -c 40 -o /tmp/r2_prof python profiles/prof_kernels.py > gpurun_out/r2_ncu_prof.log 2>&1; echo "ncu full rc=$?" >> $L
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>> $L
ncu -i /tmp/r2_prof.ncu-rep --page source --csv > /tmp/r2_prof_source.csv 2>> $L
gzip -c /tmp/r2_prof_source.csv > gpurun_out/r2_prof_source.csv.gz
ls -la /tmp/r2_prof.ncu-rep gpurun_out/ >> $L
cat $L
