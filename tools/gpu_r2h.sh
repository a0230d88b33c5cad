#!/bin/bash
# round-2 GPU call H: tile depthwise kernel at S=96 (A/B against the strip kernel), tests that cover it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_16bit.py tests/test_gpu_parity.py -q -x -k "s96 or known_answer or first_generation or every_layer or variant_b" > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2h_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2h_steps_96.txt > /dev/null 2> gpurun_out/r2h_steps_96.err
EMD_DISABLE_DW_TILE=1 timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2h_steps_96_strip.txt > /dev/null 2>&1
tail -1 gpurun_out/r2h_steps_96.txt; tail -1 gpurun_out/r2h_steps_96_strip.txt
grep -E "^deconv2_0:dw|^cnn2_last:dw|^mid5_1:dw|^cnn3_last:dw" gpurun_out/r2h_steps_96.txt gpurun_out/r2h_steps_96_strip.txt
