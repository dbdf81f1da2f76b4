set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q --durations=8 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2m_bench_cfg4.json 2> gpurun_out/r2m_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2m_bench_cfg4.err
