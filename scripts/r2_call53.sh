#!/bin/bash
# round 2, GPU call 53: k_search with the L1 prefetch of the match child's index blocks (default build) against without
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 bash scripts/ab.sh base nopf base > gpurun_out/r2c53_ab.log 2>&1
cat gpurun_out/r2c53_ab.log
