#!/bin/bash
# round 2, GPU call 51: the request-lean k_search with 7 and 8 resident blocks per SM (72 / 64 registers)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 bash scripts/ab.sh base minb7 minb8 > gpurun_out/r2c51_ab.log 2>&1
cat gpurun_out/r2c51_ab.log
