#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/fork_probe.py 2>&1 | tail -20 | cut -c1-260
