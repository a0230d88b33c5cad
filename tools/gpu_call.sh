#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
bash tools/gpu_validate.sh r2p
timeout 120 python tools/ubench/cublas_trunk.py > gpurun_out/r2p_cublas.txt 2>&1; cat gpurun_out/r2p_cublas.txt
bash tools/gpu_ab.sh r2p_ab "EMD_DW_STAGES=6" "EMD_DW_STAGES=8" "EMD_DW_SB=3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dw_tma_kernel -s 1 -c 1 -f -o gpurun_out/r2p_dwtma python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2p_ncu_dwtma.log 2>&1; echo "ncu dw rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 1 -c 1 -f -o gpurun_out/r2p_mid python tools/run_layer.py --layer mid5_1 --n 32 --mode fp16 > gpurun_out/r2p_ncu_mid.log 2>&1; echo "ncu mid rc=$?"
