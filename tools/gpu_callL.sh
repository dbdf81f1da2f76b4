set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python __graft_entry__.py 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r2l_dist_check_2gpu.log 2>&1; echo "dist check rc=$?"
tail -20 gpurun_out/r2l_dist_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2l_bench_cfg4_2gpu.json 2> gpurun_out/r2l_bench_cfg4_2gpu.err; echo "bench 2gpu rc=$?"
tail -5 gpurun_out/r2l_bench_cfg4_2gpu.err; cat gpurun_out/r2l_bench_cfg4_2gpu.json | cut -c1-1500
