#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2ae_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2ae_d00 python tools/run_layer.py --layer deconv0_0 --n 8 --mode fp16 > gpurun_out/r2ae_ncu_d00.log 2>&1; echo "ncu d00 rc=$?"
timeout 120 python tools/run_layer.py --layer mid5_1 --crop 96 --n 1024 --mode fp16 > gpurun_out/r2ae_plain96.log 2>&1; echo "plain96 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_reg_kernel -s 1 -c 1 -f -o gpurun_out/r2ae_dwreg python tools/run_layer.py --layer mid5_1 --crop 96 --n 1024 --mode fp16 > gpurun_out/r2ae_ncu_dwreg.log 2>&1; echo "ncu dwreg rc=$?"
