set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
python tools/nn_only_probe.py cfg4 > gpurun_out/r2i_nn_only.json 2> gpurun_out/r2i_nn_only.err; echo "probe rc=$?"; cat gpurun_out/r2i_nn_only.json
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2i_bench_cfg4.json 2> gpurun_out/r2i_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2i_bench_cfg4.err
