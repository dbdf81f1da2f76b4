set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --durations=5 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2j_bench_cfg4.json 2> gpurun_out/r2j_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2j_bench_cfg4.err
VP_SEARCH_OCC8=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2j_bench_cfg4_occ8.json 2> gpurun_out/r2j_bench_cfg4_occ8.err; echo "bench occ8 rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/r2j_bench_cfg3.json 2> gpurun_out/r2j_bench_cfg3.err; echo "bench3 rc=$?"
python tools/nn_only_probe.py cfg4 > gpurun_out/r2j_nn_only.json 2> gpurun_out/r2j_nn_only.err; cat gpurun_out/r2j_nn_only.json
