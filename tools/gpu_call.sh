#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py > gpurun_out/r2ak_rest.log 2>&1; echo "rest rc=$?"; tail -2 gpurun_out/r2ak_rest.log
