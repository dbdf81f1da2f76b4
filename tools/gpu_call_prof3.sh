set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_launchrun.json 2> gpurun_out/r2f_launchrun.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_search|k_bin|k_cell|k_fft|k_scan|k_plane' -c 400 --csv --log-file gpurun_out/r2f_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_launches_ncu.log 2>&1; echo "launch list rc=$?"
grep -c . gpurun_out/r2f_launches_cfg4.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_bin_hist|k_bin_scatter|k_cell_count|k_cell_place|k_search_brick|k_search_block4|k_search_exact|k_fft_z|k_fft_y|k_fft_x_pow|k_bin_tiles|k_scan_top' -c 20 -o gpurun_out/r2f_prof python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2f_prof.ncu-rep
ncu -i gpurun_out/r2f_prof.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_raw.csv 2>/dev/null; wc -c gpurun_out/r2f_ncu_raw.csv
