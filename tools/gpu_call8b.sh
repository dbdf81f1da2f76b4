set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv
python __graft_entry__.py 2>&1 | tail -2
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530+n)) "$@"; }
run 8 tests/dist_gpu_check.py > gpurun_out/r2_final_dist_check_8gpu.log 2>&1; echo "dist check rc=$?"; tail -8 gpurun_out/r2_final_dist_check_8gpu.log
run 8 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_final_bench_cfg4_8gpu.json 2> gpurun_out/r2_final_bench_cfg4_8gpu.err; echo "cfg4x8 rc=$?"; cut -c1-400 gpurun_out/r2_final_bench_cfg4_8gpu.json
run 8 bench.py --gpus 8 --steps 2 --warmup 1 --workload cfg5 --no-cpu --no-e2e > gpurun_out/r2_final_bench_cfg5_8gpu.json 2> gpurun_out/r2_final_bench_cfg5_8gpu.err; echo "cfg5 rc=$?"; cut -c1-400 gpurun_out/r2_final_bench_cfg5_8gpu.json
run 4 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_final_bench_cfg4_4gpu.json 2> gpurun_out/r2_final_bench_cfg4_4gpu.err; echo "cfg4x4 rc=$?"
