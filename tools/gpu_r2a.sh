#!/bin/bash
# round-2 GPU call A: full GPU suite, smoke, bench in both 16-bit modes, per-step profiles, pipe microbenchmark
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_fp16.json 2> gpurun_out/r2a_bench_fp16.err; echo "bench fp16 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --mode bf16 --no-configs --no-cpu-baseline > gpurun_out/r2a_bench_bf16.json 2> gpurun_out/r2a_bench_bf16.err; echo "bench bf16 rc=$?"
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2a_steps_fp16.txt > /dev/null 2> gpurun_out/r2a_steps_fp16.err
timeout 300 python tools/profile_steps.py --mode bf16 --out gpurun_out/r2a_steps_bf16.txt > /dev/null 2> gpurun_out/r2a_steps_bf16.err
timeout 120 tools/ubench/pipes > gpurun_out/r2a_pipes.txt 2>&1
cut -c1-600 gpurun_out/r2a_bench_fp16.json; tail -3 gpurun_out/r2a_bench_fp16.err
