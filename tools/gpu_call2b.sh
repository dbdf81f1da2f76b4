set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python __graft_entry__.py 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_gpu_check.py > gpurun_out/r2b_dist_check_2gpu.log 2>&1; echo "dist check rc=$?"; tail -8 gpurun_out/r2b_dist_check_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_bench_cfg4_2gpu.json 2> gpurun_out/r2b_bench_cfg4_2gpu.err; echo "cfg4x2 rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r2b_bench_cfg4_2gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['result']['nsample_crc32'], d['result']['psum']['velocity']['sum'], d.get('nvlink'))
for k,v in d['stages'].items(): print(k, v['ms_per_step'])
print(d['dist_phases_ms_last_step'])
P
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2b_bench_cfg4_1gpu.json 2>/dev/null; python - <<P
import json
d=json.loads(open('gpurun_out/r2b_bench_cfg4_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['result']['nsample_crc32']); print({k:v['ms_per_step'] for k,v in d['stages'].items() if k.startswith('k1a') or k.startswith('k1b')})
P
