set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_cfg4_1gpu_final.json 2> gpurun_out/r2_bench_cfg4_1gpu_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_cfg4_1gpu_final.err
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref rc=$?"
for wl in cfg1 cfg2 cfg3; do
timeout 900 python bench.py --workload $wl > gpurun_out/r2_bench_${wl}_final.json 2> gpurun_out/r2_bench_${wl}_final.err; echo "bench $wl rc=$?"
done
timeout 600 python tools/bench_deposit.py > gpurun_out/r2_bench_deposit.jsonl 2> gpurun_out/r2_bench_deposit.err; echo "deposit rc=$?"; cat gpurun_out/r2_bench_deposit.jsonl
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_bench_cfg4_launchrun.json 2> gpurun_out/r2_launchrun.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg4.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_launches_ncu.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_bin_hist|k_bin_scatter|k_cell_count|k_cell_place|k_search_brick|k_search_block4|k_search_exact|k_fft_z|k_fft_y|k_fft_x_pow|k_bin_tiles' -c 17 -o gpurun_out/r2_final_prof python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_final_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -20
