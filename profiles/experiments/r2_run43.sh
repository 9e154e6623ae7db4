cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp43.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp43_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp43_tests.log >> $L
for m in "30000000 20" "40000000 15" "24000000 26" "125000000 5"; do
  timeout 300 python profiles/experiments/r2_elem_on_poisson5.py $m 2>&1 | grep "sum\|argmax" >> $L
done
cat $L
