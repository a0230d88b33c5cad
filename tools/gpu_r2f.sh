#!/bin/bash
# round-2 GPU call F: single polling epilogue warp, strip depthwise v2; full suite, profiles (512 / 96), bench, ncu launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2f_tests.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/r2f_steps.txt > /dev/null 2> gpurun_out/r2f_steps.err; echo "prof rc=$?"
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/r2f_steps_96.txt > /dev/null 2> gpurun_out/r2f_steps_96.err
tail -1 gpurun_out/r2f_steps.txt; tail -1 gpurun_out/r2f_steps_96.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
cut -c1-250 gpurun_out/r2f_bench.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2f_bench_short.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r2f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
