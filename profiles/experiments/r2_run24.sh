cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp24.log
python profiles/prof_dist.py geometric_tail > $L 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_dist_launches.csv python profiles/prof_dist.py geometric_tail > gpurun_out/r2_dist_ncu.log 2>&1
cat $L
