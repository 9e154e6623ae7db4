cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp41.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp41_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp41_tests.log >> $L
for m in "60000000 10" "75000000 8" "46000000 13" "40000000 15"; do
  timeout 300 python profiles/experiments/r2_elem_on_poisson5.py $m 2>&1 | grep "mean>7" | grep "sum\|argmax" >> $L
done
echo "== forced element tiles (mean>0) for comparison at 13" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 46000000 13 2>&1 | grep "mean>0" | grep "sum\|argmax" >> $L
cat $L
