#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2ad_bench_8gpu.json 2> gpurun_out/r2ad_bench_8gpu.err; echo "8gpu rc=$?"
cut -c1-300 gpurun_out/r2ad_bench_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2ad_bench_8gpu_ref.json 2> gpurun_out/r2ad_bench_8gpu_ref.err; echo "8gpu ref rc=$?"
cut -c1-200 gpurun_out/r2ad_bench_8gpu_ref.json
