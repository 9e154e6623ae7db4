cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp28.log
: > $L
BASE=$PWD/awkward-0.x_b200/lib/libjagged_base.so
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp28_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp28_tests.log >> $L
echo "== optable new" >> $L
timeout 300 python profiles/experiments/r2_optable.py '_select' --dist >> $L 2>&1; echo "rc=$?" >> $L
echo "== optable base" >> $L
JAGGED_B200_LIB=$BASE timeout 300 python profiles/experiments/r2_optable.py '_select' --dist >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp28_bench_new.json 2> gpurun_out/r2_exp28_bench_new.err; echo "bench new rc=$?" >> $L
JAGGED_B200_LIB=$BASE timeout 300 python bench.py --steps 20 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_exp28_bench_base.json 2> gpurun_out/r2_exp28_bench_base.err; echo "bench base rc=$?" >> $L
python - >> $L <<'PY'
import json
for w in ("new", "base"):
    try:
        d=json.loads(open('gpurun_out/r2_exp28_bench_%s.json' % w).read().strip().splitlines()[-1])
        print(w, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['totals_check'], d['roofline']['ms'], d['roofline']['frac'])
    except Exception as e:
        print(w, "failed", e)
PY
timeout 300 python profiles/experiments/r2_pyprofile_sweep.py > gpurun_out/r2_exp28_pyprofile.txt 2>&1; head -12 gpurun_out/r2_exp28_pyprofile.txt >> $L
cat $L
