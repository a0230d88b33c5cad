#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2q_tests.log
bash tools/gpu_ab.sh r2q_ab "EMD_DISABLE_PAD_PITCH=1" "EMD_DW_REG_ALL=1" "EMD_DW_STAGES=8"
for i in 1 2; do
  timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2q_96_base_$i.txt > /dev/null 2>&1
  EMD_DISABLE_DW_REG=1 timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2q_96_noreg_$i.txt > /dev/null 2>&1
  echo "96 rep $i: base $(tail -1 gpurun_out/r2q_96_base_$i.txt | cut -c24-34) | no dw_reg $(tail -1 gpurun_out/r2q_96_noreg_$i.txt | cut -c24-34)"
done
