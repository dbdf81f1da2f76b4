set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "slab" > gpurun_out/r2z2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2z2_pytest.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_bench_cfg4_launchrun.json 2> gpurun_out/r2_launchrun.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_search|k_bin|k_cell|k_fft|k_scan|k_plane' -c 400 --csv --log-file gpurun_out/r2_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_launches_ncu.log 2>&1; echo "launch list rc=$?"
grep -c . gpurun_out/r2_launches_cfg4.csv
