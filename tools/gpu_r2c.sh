#!/bin/bash
# round-2 GPU call C: flexible tile geometry (small maps), pipelined fused trunk; full suite, S=96 profile, A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r2c_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2c_steps_fuse.txt > /dev/null 2> gpurun_out/r2c_steps_fuse.err; echo "prof rc=$?"
EMD_DISABLE_TRUNK_FUSE=1 timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2c_steps_nofuse.txt > /dev/null 2>&1
EMD_DISABLE_TRUNK_FUSE=1 timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2c_steps_96.txt > /dev/null 2> gpurun_out/r2c_steps_96.err
EMD_DISABLE_TRUNK_FUSE=1 EMD_DISABLE_FUSED=1 timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2c_steps_96_gen1.txt > /dev/null 2>&1
tail -1 gpurun_out/r2c_steps_fuse.txt; tail -1 gpurun_out/r2c_steps_nofuse.txt; tail -1 gpurun_out/r2c_steps_96.txt; tail -1 gpurun_out/r2c_steps_96_gen1.txt
grep -E "^mid5_1|^mid5_2|^cnn3_last|^cnn3 " gpurun_out/r2c_steps_fuse.txt gpurun_out/r2c_steps_nofuse.txt
