#!/bin/bash
# round 2, GPU call 45: k_search with the split pop (default build) against the one-trip pop
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base nosplit base nosplit > gpurun_out/r2c45_ab.log 2>&1
cat gpurun_out/r2c45_ab.log
