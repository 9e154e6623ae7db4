cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp33.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp33_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -5 gpurun_out/r2_exp33_tests.log >> $L
timeout 600 python profiles/experiments/r2_geometric_parts.py >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python profiles/experiments/r2_optable.py 'segreduce_sum_f32|segargmax_dense_tiles_f32|compare_select_f32|mask_select_f32' >> $L 2>&1; echo "rc=$?" >> $L
cat $L
