set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "not cfg2_full and not 500 and not N1000 and not 2048 and not parseval" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_pytest.log
run() { tag=$1; wl=$2; shift; shift; env "$@" timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2x_bench_$tag.json 2> gpurun_out/r2x_bench_$tag.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2x_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', d['ms_per_step'], d['result']['nsample_crc32'], {k:round(v['ms_per_step'],3) for k,v in d['stages'].items() if k.startswith('k1')})
P
}
run cfg4_wc1b cfg4 VP_BIN_WC=1
run cfg2_wc1b cfg2 VP_BIN_WC=1
