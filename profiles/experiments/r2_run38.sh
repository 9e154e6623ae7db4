cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_exp38.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_exp38_tests.log 2>&1; echo "tests rc=$?" >> $L
tail -3 gpurun_out/r2_exp38_tests.log >> $L
timeout 300 python profiles/experiments/r2_optable.py 'xxxx' --dist >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 60000000 10 2>&1 | grep "mean>7" >> $L
timeout 300 python profiles/experiments/r2_elem_on_poisson5.py 30000000 20 2>&1 | grep "mean>7" >> $L
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "reference rc=$?" >> $L
cat $L
