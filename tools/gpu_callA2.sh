set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a2_bench_$tag.json 2> gpurun_out/r2a2_bench_$tag.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2a2_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', d['ms_per_step'], d['result']['nsample_crc32'], {k:v['ms_per_step'] for k,v in d['stages'].items() if k.startswith('k1g') or k.startswith('k1a')})
P
}
run occ6 VP_B4_OCC=6
run occ5 VP_B4_OCC=5
run occ4 VP_B4_OCC=4
