#!/bin/bash
# round-2 GPU call L: async forward pipeline (tests + bench), stem tweak
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -s > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/r2l_bench.json; tail -3 gpurun_out/r2l_bench.err
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2l_steps.txt > /dev/null 2> gpurun_out/r2l_steps.err
tail -1 gpurun_out/r2l_steps.txt; grep -E "^cnn0 |^residual0 " gpurun_out/r2l_steps.txt
