#!/bin/bash
# round-2 GPU call E: fused 728-wide blocks with a TMA halo ring (EMD_TRUNK_FUSE=1); ncu of the strip depthwise kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shape.py -q -s -k "fused_trunk or forced_pair or pair_kernels" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2e_tests.log
EMD_TRUNK_FUSE=1 timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2e_steps_fuse.txt > /dev/null 2> gpurun_out/r2e_steps_fuse.err; echo "prof rc=$?"
tail -1 gpurun_out/r2e_steps_fuse.txt; grep -E "^mid5_1|^mid5_2|^cnn3_last|^cnn3 " gpurun_out/r2e_steps_fuse.txt
EMD_TRUNK_FUSE=1 timeout 400 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/r2e_bench_fuse.json 2> gpurun_out/r2e_bench_fuse.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r2e_bench_fuse.json
timeout 200 python tools/run_layer.py --layer deconv2_0 --crop 96 --n 1024 --mode fp16 > gpurun_out/r2e_plain_strip.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_strip -s 1 -c 1 -f -o gpurun_out/r2e_strip python tools/run_layer.py --layer deconv2_0 --crop 96 --n 1024 --mode fp16 > gpurun_out/r2e_ncu_strip.log 2>&1; echo "ncu strip rc=$?"
EMD_TRUNK_FUSE=1 timeout 120 python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2e_plain_mid.log 2>&1 && \
EMD_TRUNK_FUSE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2e_mid python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2e_ncu_mid.log 2>&1; echo "ncu mid rc=$?"
