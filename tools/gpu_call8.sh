set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv
python __graft_entry__.py 2>&1 | tail -2
nvl() { nvidia-smi nvlink -gt d -i 0 > gpurun_out/$1 2>&1; }
nvl r2_nvlink_0_before_cfg5.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 1 --workload cfg5 --no-cpu --no-e2e > gpurun_out/r2_bench_cfg5_8gpu.json 2> gpurun_out/r2_bench_cfg5_8gpu.err; echo "cfg5 rc=$?"
nvl r2_nvlink_0_after_cfg5.txt
tail -5 gpurun_out/r2_bench_cfg5_8gpu.err; cut -c1-1200 gpurun_out/r2_bench_cfg5_8gpu.json
nvl r2_nvlink_0_before_cfg4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_cfg4_8gpu.json 2> gpurun_out/r2_bench_cfg4_8gpu.err; echo "cfg4x8 rc=$?"
nvl r2_nvlink_0_after_cfg4.txt
tail -5 gpurun_out/r2_bench_cfg4_8gpu.err; cut -c1-600 gpurun_out/r2_bench_cfg4_8gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tests/dist_gpu_check.py > gpurun_out/r2_dist_check_8gpu.log 2>&1; echo "dist check rc=$?"
tail -14 gpurun_out/r2_dist_check_8gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_cfg4_4gpu.json 2> gpurun_out/r2_bench_cfg4_4gpu.err; echo "cfg4x4 rc=$?"
