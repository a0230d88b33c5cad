#!/bin/bash
# Sweep tuning switches: each argument is one "VAR=val VAR=val" setting; prints total and the listed layers for each.
#   tools/sweep_env.sh "EMD_DW_SH=3" "EMD_DW_SA=3 EMD_DW_SH=5"
LAYERS="${LAYERS:-cnn0_last cnn1 cnn1_last cnn2 deconv2_0 deconv2_1 deconv1_0 deconv1_1 deconv0_0 deconv0_1 TOTAL}"
for setting in "$@"; do
  echo "== $setting"
  env $setting timeout 120 python tools/profile_steps.py --batch 32 > /tmp/sweep.txt 2>&1 || { tail -3 /tmp/sweep.txt; continue; }
  for l in $LAYERS; do grep -E "^$l " /tmp/sweep.txt | awk '{printf "   %-14s %s\n", $1, $2}'; done
done
