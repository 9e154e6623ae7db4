cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp31.log
: > $L
BASE=$PWD/awkward-0.x_b200/lib/libjagged_base.so
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp31_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp31_tests.log >> $L
for w in new base; do
  echo "== $w" >> $L
  if [ $w = base ]; then export JAGGED_B200_LIB=$BASE; fi
  timeout 300 python profiles/experiments/r2_optable.py 'parents|localindex|broadcast|add_flat|index_gather|distincts|segment' --dist >> $L 2>&1; echo "rc=$?" >> $L
  timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 60000000 10 2>&1 | grep "mean>7" >> $L
  timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 30000000 20 2>&1 | grep "mean>7" >> $L
done
cat $L
