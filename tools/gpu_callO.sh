set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2o_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2o_bench_cfg4.json 2> gpurun_out/r2o_bench_cfg4.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r2o_bench_cfg4.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2o_bench_cfg4.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['nn_stats'], d['result']['nsample_crc32'], d['result']['psum']['velocity']['sum'])
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
P
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2o_bench_cfg3.json 2> gpurun_out/r2o_bench_cfg3.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2o_bench_cfg3.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['nn_stats'], d['result']['nsample_crc32'])
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
P
