set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nn or golden or library or cfg or slab or spot or script or pk_fields or parseval or full_size" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2v_pytest.log
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2v_bench_$tag.json 2> gpurun_out/r2v_bench_$tag.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2v_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', d['ms_per_step'], d['nn_stats'], d['result']['nsample_crc32'], d['result']['psum']['velocity']['sum'])
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
P
}
run base A=1
run s512 VP_SCATTER_512=1
run noswz VP_X_NOSWIZZLE=1
