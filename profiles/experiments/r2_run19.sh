cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp19.log
CMD="python profiles/experiments/r2_optable.py pairs_values|distincts_values --events 20000000 --comb-events 20000000 --kernel-reps 2"
timeout 300 $CMD > $L 2>&1 || { cat $L; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'comb_gather2_kernel' -s 2 -c 5 -f -o /tmp/g2 $CMD > gpurun_out/r2_ncu_g2.log 2>&1
ncu -i /tmp/g2.ncu-rep --page raw --csv > gpurun_out/r2_g2_raw.csv 2>/dev/null
ncu -i /tmp/g2.ncu-rep --page source --csv > gpurun_out/r2_g2_source.csv 2>/dev/null
ls -la gpurun_out/r2_g2* 
cat $L
