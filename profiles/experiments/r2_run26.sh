cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp26.log
: > $L
timeout 200 python profiles/experiments/r2_optable.py '_select' --dist >> $L 2>&1; echo "optable rc=$?" >> $L
JG_SELECT_SINGLE=1 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp26_tests.log 2>&1; echo "tests(single) rc=$?" >> $L
tail -3 gpurun_out/r2_exp26_tests.log >> $L
JG_SELECT_SINGLE=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp26_bench.json 2> gpurun_out/r2_exp26_bench.err; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r2_exp26_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['totals_check'])
PY
cat $L
