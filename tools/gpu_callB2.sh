set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -10 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench_cfg4_1gpu_final.json 2> gpurun_out/r2_bench_cfg4_1gpu_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_cfg4_1gpu_final.json
