cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp29.log
: > $L
BASE=$PWD/awkward-0.x_b200/lib/libjagged_base.so
timeout 600 python -m pytest tests -m gpu -q -x -k "argm or selection_row_edges or golden or long_event" > gpurun_out/r2_exp29_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp29_tests.log >> $L
echo "== optable new" >> $L
timeout 300 python profiles/experiments/r2_optable.py 'segargm' >> $L 2>&1; echo "rc=$?" >> $L
echo "== optable base" >> $L
JAGGED_B200_LIB=$BASE timeout 300 python profiles/experiments/r2_optable.py 'segargm' >> $L 2>&1; echo "rc=$?" >> $L
echo "== elem variants forced" >> $L
timeout 400 python profiles/experiments/r2_elem_on_short.py >> $L 2>&1; echo "rc=$?" >> $L
cat $L
