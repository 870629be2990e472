#!/bin/bash
# round 2, GPU call 44: k_search / k_width on C4 with .L2::64B on the index loads
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base l2_64 base l2_64 > gpurun_out/r2c44_ab.log 2>&1
cat gpurun_out/r2c44_ab.log
