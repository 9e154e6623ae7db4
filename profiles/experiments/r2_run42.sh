cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp42.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp42_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp42_tests.log >> $L
for m in "60000000 10" "75000000 8" "46000000 13" "125000000 5"; do
  timeout 300 python profiles/experiments/r2_elem_on_poisson5.py $m 2>&1 | grep "mean>7" >> $L
done
timeout 300 python profiles/experiments/r2_optable.py 'parents|localindex|broadcast' >> $L 2>&1
cat $L
