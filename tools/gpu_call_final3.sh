set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -k "interleaved_rows or slab_bucket_kernel" > gpurun_out/r2f_pytest_rows.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2f_pytest_rows.log
timeout 60 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_launchrun.json 2> gpurun_out/r2f_launchrun.err && \
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_search|k_bin|k_cell|k_fft|k_scan|k_plane' -c 400 --csv --log-file gpurun_out/r2f_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_launches_ncu.log 2>&1; echo "launch list rc=$?"
grep -c . gpurun_out/r2f_launches_cfg4.csv
