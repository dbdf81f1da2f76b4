set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export VP_NO_CLUSTERS=1
python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g_plain.json 2> gpurun_out/r2g_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_search_rows|k_bin_scatter|k_cell_place|k_fft_x_pow|k_bin_tiles|k_cell_count' -c 7 -o gpurun_out/r2g_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2g_ncu.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r2g_ncu.log
ls -la gpurun_out/
