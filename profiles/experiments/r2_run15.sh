cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp15.log
timeout 300 python profiles/prof_comb.py > $L 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'comb_.*' -c 12 -f -o /tmp/comb python profiles/prof_comb.py > gpurun_out/r2_ncu_comb.log 2>&1
ncu -i /tmp/comb.ncu-rep --page raw --csv > gpurun_out/r2_comb_raw.csv 2>/dev/null
ncu -i /tmp/comb.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_comb_source.csv 2>/dev/null
ls -la /tmp/comb.ncu-rep gpurun_out/ | tail -5
cat $L
