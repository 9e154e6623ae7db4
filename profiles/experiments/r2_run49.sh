cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp49.log
: > $L
JG_SELECT_RANGES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "golden or differential or selection_row or tiles_that or c1_full or unaligned or edge_layouts" > gpurun_out/r2_exp49_tests.log 2>&1; echo "tests(ranges) rc=$?" >> $L
tail -3 gpurun_out/r2_exp49_tests.log >> $L
echo "== three launches" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 125000000 5 2>&1 | grep "mean>7" | grep "a\[a" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 200000000 2 2>&1 | grep "mean>7" | grep "a\[a" >> $L
echo "== ranges" >> $L
JG_SELECT_RANGES=1 timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 125000000 5 2>&1 | grep "mean>7" | grep "a\[a" >> $L
JG_SELECT_RANGES=1 timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 200000000 2 2>&1 | grep "mean>7" | grep "a\[a" >> $L
JG_SELECT_RANGES=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp49_bench.json 2> gpurun_out/r2_exp49_bench.err; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r2_exp49_bench.json').read().strip().splitlines()[-1])
print('ranges sweep', d['ms_per_step'], d['value'], d['totals_check'])
PY
cat $L
