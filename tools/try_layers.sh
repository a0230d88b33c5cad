#!/bin/bash
# Run each listed layer alone (tools/run_layer.py) under a short timeout; prints ok / rc per layer.
N="${N:-8}"
for l in "$@"; do
  timeout 25 python tools/run_layer.py --layer $l --n $N > /tmp/try_$l.log 2>&1; rc=$?
  echo "$l rc=$rc $(tail -1 /tmp/try_$l.log | cut -c1-150)"
done
