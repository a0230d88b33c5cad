#!/bin/bash
# round-2 GPU call M: math-warp addressing / smem weights: dw-layer tests + A/B profile against the previous build on one box
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_bench_shape.py tests/test_gpu_16bit.py -q -x > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2m_tests.log
for i in 1 2; do
  timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2m_steps_new_$i.txt > /dev/null 2>&1
  EMD_DISABLE_DW_COLS=1 timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2m_steps_oldmap_$i.txt > /dev/null 2>&1
  echo "rep $i: new $(tail -1 gpurun_out/r2m_steps_new_$i.txt | cut -c1-50) | first mapping $(tail -1 gpurun_out/r2m_steps_oldmap_$i.txt | cut -c1-50)"
done
grep -E "^deconv0_0|^deconv1_0|^cnn0_last|^deconv0_1|^deconv2_0|^cnn2_last|^deconv2_1|^cnn1 " gpurun_out/r2m_steps_new_1.txt
