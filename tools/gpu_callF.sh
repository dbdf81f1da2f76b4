set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --durations=5 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2f_bench_cfg4.json 2> gpurun_out/r2f_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2f_bench_cfg4.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err; echo "bench3 rc=$?"
VP_NO_CLUSTERS=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_noclu.json 2> gpurun_out/r2f_bench_cfg4_noclu.err; echo "bench noclu rc=$?"
VP_INLINE_B=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_inlineb.json 2> gpurun_out/r2f_bench_cfg4_inlineb.err; echo "bench inlineb rc=$?"
VP_SEARCH_BRICK=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_brick.json 2> gpurun_out/r2f_bench_cfg4_brick.err; echo "bench brick rc=$?"
