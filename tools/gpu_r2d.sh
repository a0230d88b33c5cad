#!/bin/bash
# round-2 GPU call D: mbarrier wait limit, strip depthwise for small maps; full suite + profiles + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2d_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2d_steps.txt > /dev/null 2> gpurun_out/r2d_steps.err; echo "prof rc=$?"
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2d_steps_96.txt > /dev/null 2> gpurun_out/r2d_steps_96.err
tail -1 gpurun_out/r2d_steps.txt; tail -1 gpurun_out/r2d_steps_96.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/r2d_bench.json
timeout 120 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2d_plain_d00.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2d_d00 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2d_ncu_d00.log 2>&1; echo "ncu d00 rc=$?"
