cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp25.log
: > $L
timeout 300 python profiles/experiments/r2_pyprofile_sweep.py 2>&1 | head -3 >> $L
timeout 900 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp25_bench.json 2> gpurun_out/r2_exp25_bench.err; echo "bench rc=$?" >> $L
python - >> $L <<'PY'
import json
d=json.loads(open('gpurun_out/r2_exp25_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['totals_check'])
PY
cat $L
