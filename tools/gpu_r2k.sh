#!/bin/bash
# round-2 GPU call K: full GPU suite, smoke, per-step profiles (512 / 96), bench (both arms) on the committed build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2k_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2k_smoke.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2k_steps.txt > /dev/null 2> gpurun_out/r2k_steps.err
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2k_steps_96.txt > /dev/null 2> gpurun_out/r2k_steps_96.err
tail -1 gpurun_out/r2k_steps.txt; tail -1 gpurun_out/r2k_steps_96.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/r2k_bench.json
