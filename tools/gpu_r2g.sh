#!/bin/bash
# round-2 GPU call G: A/B of the single polling epilogue warp on ONE box (3 alternating repeats), persistent strip depthwise at S=96
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2g_steps_poll1_$i.txt > /dev/null 2>&1
  EMD_DISABLE_EPI_POLL1=1 timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2g_steps_pollall_$i.txt > /dev/null 2>&1
  echo "rep $i: poll1 $(tail -1 gpurun_out/r2g_steps_poll1_$i.txt | cut -c1-50) | pollall $(tail -1 gpurun_out/r2g_steps_pollall_$i.txt | cut -c1-50)"
done
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2g_steps_96.txt > /dev/null 2> gpurun_out/r2g_steps_96.err
tail -1 gpurun_out/r2g_steps_96.txt
timeout 600 python -m pytest tests/test_gpu_16bit.py tests/test_gpu_parity.py -q -x -k "s96 or known_answer or first_generation or variant_b" > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_tests.log
