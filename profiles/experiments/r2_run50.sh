cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp50.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp50_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -2 gpurun_out/r2_exp50_tests.log >> $L
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> $L
timeout 300 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp50_bench.json 2> gpurun_out/r2_exp50_bench.err; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r2_exp50_bench.json').read().strip().splitlines()[-1])
print('sweep', d['ms_per_step'], d['value'], d['totals_check'], d['e2e']['ms_per_step'])
PY
cat $L
