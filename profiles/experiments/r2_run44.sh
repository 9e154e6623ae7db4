cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp44.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp44_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp44_tests.log >> $L
echo "== new" >> $L
timeout 300 python profiles/experiments/r2_addflat.py >> $L 2>&1
echo "== base" >> $L
JAGGED_B200_LIB=$PWD/awkward-0.x_b200/lib/libjagged_base.so timeout 300 python profiles/experiments/r2_addflat.py >> $L 2>&1
cat $L
