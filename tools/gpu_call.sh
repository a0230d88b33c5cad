#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== repro, 16 crops two chunks"; timeout 600 python tools/keep_repro.py 16 2>&1 | tail -4 | cut -c1-300
echo "== repro, 8 crops in a 16-crop workspace"; timeout 600 python tools/keep_repro.py 16 8 16 2>&1 | tail -4 | cut -c1-300
bash tools/gpu_validate.sh r2aa
