cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-tables > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --total-events 1e9 --steps 5 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_bench_1B_8gpu.json 2> gpurun_out/r2_bench_1B_8gpu.err; echo "rc=$?"
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_bench_1gpu_on8box.json 2> gpurun_out/r2_bench_1gpu_on8box.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_1gpu_on8box','r2_bench_4gpu','r2_bench_8gpu','r2_bench_1B_8gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ('value','ms_per_step','n_gpus','totals_check','scaling')}, d['e2e']['ms_per_step'] if d.get('e2e') else None, d.get('nccl_gather_per_event_f32',{}).get('gbs_per_rank') if d.get('nccl_gather_per_event_f32') else None)
    except Exception as e:
        print(f, 'failed', e)
PY
