#!/bin/bash
# round-2 GPU call O: stage-count experiments of the fused depthwise mode (same box, alternating)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2; do
  timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2o_base_$i.txt > /dev/null 2>&1
  EMD_DW_SA=2 EMD_DW_SB=2 EMD_DW_SH=6 timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2o_sa2sb2_$i.txt > /dev/null 2>&1
  EMD_DW_SB=2 EMD_DW_SH=6 timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2o_sb2_$i.txt > /dev/null 2>&1
  echo "rep $i: base $(tail -1 gpurun_out/r2o_base_$i.txt | cut -c24-34) | SA2 SB2 $(tail -1 gpurun_out/r2o_sa2sb2_$i.txt | cut -c24-34) | SB2 $(tail -1 gpurun_out/r2o_sb2_$i.txt | cut -c24-34)"
done
