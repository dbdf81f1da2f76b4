set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2k_bench_cfg4.json 2> gpurun_out/r2k_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2k_bench_cfg4.err
VP_SEARCH_BLOCK=256 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2k_bench_cfg4_b256.json 2> gpurun_out/r2k_bench_cfg4_b256.err; echo "bench b256 rc=$?"
VP_SEARCH_BLOCK=64 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2k_bench_cfg4_b64.json 2> gpurun_out/r2k_bench_cfg4_b64.err; echo "bench b64 rc=$?"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "nn or fused or cfg" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
