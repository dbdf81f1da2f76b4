set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 240 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -9 gpurun_out/r2f_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py > gpurun_out/r2f_bench_cfg4_1gpu.json 2> gpurun_out/r2f_bench_cfg4_1gpu.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2f_bench_cfg4_1gpu.json
