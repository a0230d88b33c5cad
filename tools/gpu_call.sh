#!/bin/bash
# scratch driver of the current GPU call (rewritten per call; the reusable pieces are gpu_validate.sh and gpu_ab.sh)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2al_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2al_smoke.log
