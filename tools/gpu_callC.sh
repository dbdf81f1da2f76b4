set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --durations=8 > gpurun_out/r2c_pytest_brick.log 2>&1; echo "pytest(brick) rc=$?"
tail -25 gpurun_out/r2c_pytest_brick.log
VP_NO_BRICK=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "nn or library or cfg or slab or script" > gpurun_out/r2c_pytest_nobrick.log 2>&1; echo "pytest(nobrick) rc=$?"
tail -8 gpurun_out/r2c_pytest_nobrick.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2c_bench_cfg4.json 2> gpurun_out/r2c_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2c_bench_cfg4.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; echo "bench3 rc=$?"
VP_NO_BRICK=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_bench_cfg4_nobrick.json 2> gpurun_out/r2c_bench_cfg4_nobrick.err; echo "bench nobrick rc=$?"
VP_BUCKET_SHIFT=20 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_bench_cfg4_bs20.json 2> gpurun_out/r2c_bench_cfg4_bs20.err; echo "bench bs20 rc=$?"
VP_BUCKET_SHIFT=15 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2c_bench_cfg4_bs15.json 2> gpurun_out/r2c_bench_cfg4_bs15.err; echo "bench bs15 rc=$?"
