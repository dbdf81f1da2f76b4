set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pk_fields or parseval or full_size or library or fft or golden" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2w_pytest.log
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2w_bench_$tag.json 2> gpurun_out/r2w_bench_$tag.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2w_bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', d['ms_per_step'], d['result']['nsample_crc32'], d['result']['psum']['velocity']['sum'])
for k,v in d['stages'].items():
    if k.startswith('k4') or k.startswith('k5'): print(k, v['ms_per_step'])
P
}
run g1024 VP_FFT_GROUP=1024
run g16 VP_FFT_GROUP=16
run g32 VP_FFT_GROUP=32
run g8 VP_FFT_GROUP=8
