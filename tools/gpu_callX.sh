set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nn or golden or library or cfg or slab or spot or script" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2x_pytest.log
for wl in cfg4 cfg3; do
timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2x_bench_$wl.json 2> gpurun_out/r2x_bench_$wl.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2x_bench_$wl.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['nn_stats'], d['result']['nsample_crc32'], d['result']['psum'])
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
P
done
