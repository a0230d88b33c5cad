#!/bin/bash
# round-2 GPU call J (2 GPUs): bench.py under torch.distributed.run -- the stream_4096 config with the cross-rank hash check
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -c 1500 gpurun_out/r2j_bench_2gpu.json; tail -5 gpurun_out/r2j_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2j_ref_2gpu.json 2> gpurun_out/r2j_ref_2gpu.err; echo "ref 2gpu rc=$?"
cut -c1-200 gpurun_out/r2j_ref_2gpu.json
