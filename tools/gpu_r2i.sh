#!/bin/bash
# round-2 GPU call I: per-tile tap skipping, tile depthwise v2; GPU suite, profiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2i_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2i_steps.txt > /dev/null 2> gpurun_out/r2i_steps.err
EMD_DISABLE_SKIP_TAPS=1 timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2i_steps_noskip.txt > /dev/null 2>&1
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2i_steps_96.txt > /dev/null 2> gpurun_out/r2i_steps_96.err
tail -1 gpurun_out/r2i_steps.txt; tail -1 gpurun_out/r2i_steps_noskip.txt; tail -1 gpurun_out/r2i_steps_96.txt
grep -E "^aspp_r|^deconv2to1|^deconv1to0" gpurun_out/r2i_steps.txt gpurun_out/r2i_steps_noskip.txt
grep -E "^deconv2_0:dw|^cnn2_last:dw|^mid5_1:dw|^cnn3_last:dw|^aspp_r" gpurun_out/r2i_steps_96.txt
