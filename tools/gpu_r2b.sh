#!/bin/bash
# round-2 GPU call B: new depthwise thread mapping + fused 728-wide blocks: correctness, A/B per-step profiles, ncu captures
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_16bit.py -q -s -x > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2b_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2b_steps_new.txt > /dev/null 2> gpurun_out/r2b_steps_new.err; echo "prof new rc=$?"
EMD_DISABLE_TRUNK_FUSE=1 timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2b_steps_nofuse.txt > /dev/null 2>&1
EMD_DISABLE_DW_COLS=1 timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2b_steps_nocols.txt > /dev/null 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_steps_new.txt; tail -1 gpurun_out/r2b_steps_nofuse.txt; tail -1 gpurun_out/r2b_steps_nocols.txt
cut -c1-300 gpurun_out/r2b_bench.json
timeout 120 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2b_plain_d00.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2b_d00 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2b_ncu_d00.log 2>&1; echo "ncu d00 rc=$?"
timeout 120 python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2b_plain_mid.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2b_mid python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2b_ncu_mid.log 2>&1; echo "ncu mid rc=$?"
