set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "nn or golden or library" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2p_bench_cfg4.json 2> gpurun_out/r2p_bench_cfg4.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2p_bench_cfg4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['nn_stats'], d['result']['nsample_crc32'], d['result']['psum']['velocity']['sum'])
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
P
ncu --set full --clock-control none --import-source on -k regex:'k_search_brick' -c 1 -o gpurun_out/r2p_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2p_ncu.log 2>&1
echo "ncu rc=$?"
