cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
nproc >> gpurun_out/r2_topo_8gpu.txt; free -g >> gpurun_out/r2_topo_8gpu.txt; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" >> gpurun_out/r2_topo_8gpu.txt
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-tables > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "weak8 rc=$?"
timeout 900 $T --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-tables --total-events 1e9 > gpurun_out/r2_bench_1B_8gpu.json 2> gpurun_out/r2_bench_1B_8gpu.err; echo "strong8 rc=$?"
timeout 900 $T --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 10 --warmup 3 --no-tables > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; echo "weak4 rc=$?"
tail -3 gpurun_out/r2_bench_8gpu.err gpurun_out/r2_bench_1B_8gpu.err gpurun_out/r2_bench_4gpu.err
head -c 2600 gpurun_out/r2_bench_8gpu.json; echo; head -c 1500 gpurun_out/r2_bench_1B_8gpu.json; echo; head -c 2600 gpurun_out/r2_bench_4gpu.json
