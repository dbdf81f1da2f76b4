set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench_cfg4_2gpu.json 2> gpurun_out/r2f_bench_cfg4_2gpu.err; echo "cfg4x2 rc=$?"
tail -3 gpurun_out/r2f_bench_cfg4_2gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_bench_cfg4_2gpu.json').read().strip().splitlines()[-1])
print('N=2', d['ms_per_step'], d['result']['nsample_crc32'], d['result'].get('nsample_closed_form_ok'), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})
P
