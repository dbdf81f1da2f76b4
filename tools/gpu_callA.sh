set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "cfg2_full or cfg3_clustered or cfg4_full or two_rank" --durations=5 > gpurun_out/r2a_pytest_new.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest_new.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err; echo "bench3 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --workload cfg1 > gpurun_out/r2a_bench_cfg1.json 2> gpurun_out/r2a_bench_cfg1.err; echo "bench1 rc=$?"
tail -3 gpurun_out/r2a_bench_cfg4.err gpurun_out/r2a_bench_cfg3.err gpurun_out/r2a_bench_cfg1.err
