cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python profiles/experiments/r2_cliffs.py > gpurun_out/r2_exp45.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exp45.log
cat gpurun_out/r2_exp45.log
