cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_2gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-tables > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
tail -5 gpurun_out/r2_bench_2gpu.err
cat gpurun_out/r2_bench_2gpu.json | head -c 3000
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-tables --no-cpu > gpurun_out/r2_bench_1gpu_notables.json 2> gpurun_out/r2_bench_1gpu_notables.err; echo "rc=$?"
cat gpurun_out/r2_bench_1gpu_notables.json | head -c 2500
