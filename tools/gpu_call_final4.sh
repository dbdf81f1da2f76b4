set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section ComputeWorkloadAnalysis"
timeout 100 ncu $SEC --clock-control none -k regex:k_fft_x_pow_tma -s 1 -c 1 -o /tmp/r2f_xtma python tools/pk_only_probe.py 1024 1 1 > gpurun_out/r2f_ncu_xtma.log 2>&1; echo "ncu xtma rc=$?"
ncu -i /tmp/r2f_xtma.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_k_fft_x_pow_tma_raw.csv 2>/dev/null; wc -c gpurun_out/r2f_ncu_k_fft_x_pow_tma_raw.csv
timeout 100 ncu $SEC --clock-control none -k regex:k_bin_scatter_wc -s 1 -c 1 -o /tmp/r2f_wc python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2f_ncu_wc.log 2>&1; echo "ncu wc rc=$?"
ncu -i /tmp/r2f_wc.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_k_bin_scatter_wc_cfg3_raw.csv 2>/dev/null; wc -c gpurun_out/r2f_ncu_k_bin_scatter_wc_cfg3_raw.csv
