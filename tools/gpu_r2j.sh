#!/bin/bash
# round-2 GPU call J (2 GPUs): bench.py under torch.distributed.run -- the stream_4096 config with the cross-rank hash check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -c 1200 gpurun_out/r2j_bench_2gpu.json; tail -3 gpurun_out/r2j_bench_2gpu.err
