#!/bin/bash
# Same-box A/B of tuning options on the per-step profile (boxes differ by 3-5 %, so only same-box pairs mean anything):
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh tag "EMD_DW_SB=3" "EMD_DW_SB=2 EMD_DW_SH=6"'
# runs base and every given environment alternately, twice, and prints the whole-step times side by side.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=$1; shift
mkdir -p gpurun_out
for i in 1 2; do
  timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/${T}_base_$i.txt > /dev/null 2>&1
  line="rep $i: base $(tail -1 gpurun_out/${T}_base_$i.txt | cut -c24-34)"
  k=0
  for envs in "$@"; do
    k=$((k+1))
    env $envs timeout 200 python tools/profile_steps.py --mode fp16 --out gpurun_out/${T}_v${k}_$i.txt > /dev/null 2>&1
    line="$line | [$envs] $(tail -1 gpurun_out/${T}_v${k}_$i.txt | cut -c24-34)"
  done
  echo "$line"
done
