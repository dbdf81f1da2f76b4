set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2n_pytest.log
timeout 900 python bench.py > gpurun_out/r2n_bench_cfg4.json 2> gpurun_out/r2n_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2n_bench_cfg4.err
timeout 600 python bench.py --impl reference > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; echo "ref rc=$?"
