#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2ai_bench_4gpu.json 2> gpurun_out/r2ai_bench_4gpu.err; echo "4gpu rc=$?"
cut -c1-300 gpurun_out/r2ai_bench_4gpu.json
