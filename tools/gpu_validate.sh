#!/bin/bash
# One GPU call that validates the committed build end to end (run it through gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh [tag]'
# full GPU suite, smoke(), per-step profiles (512^2 x 32 and 96^2 x 4096), both bench arms, and -- only after the bench has
# exited 0 without a profiler -- the ncu launch list of the same bench command.  Everything lands in gpurun_out/<tag>_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-val}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 300 python tools/profile_steps.py --mode fp16 --out gpurun_out/${T}_steps.txt > /dev/null 2> gpurun_out/${T}_steps.err
timeout 300 python tools/profile_steps.py --mode fp16 --batch 4096 --crop 96 --out gpurun_out/${T}_steps_96.txt > /dev/null 2> gpurun_out/${T}_steps_96.err
tail -1 gpurun_out/${T}_steps.txt; tail -1 gpurun_out/${T}_steps_96.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; rc=$?; echo "bench rc=$rc"
cut -c1-400 gpurun_out/${T}_bench.json
if [ $rc -eq 0 ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
