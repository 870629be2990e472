#!/bin/bash
# round 2, GPU call 38: cold-process SE run with BGZF input -- allocation / lane / stage traces, and the host-inflate variant
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python scripts/bam2bam_bench.py --mode se --reads 4000000 --len 76 --genome-bp 100000000 --cpu-sample 5000 \
  --log-dir $O/r2c38_logs --gpu-env trace:BWAGPU_TRACE=1 --gpu-env hostinflate:BWAGPU_HOST_INFLATE=1 --gpu-env noramp:BWAGPU_BATCH_RAMP=0 \
  --out $O/r2c38_b2b_se76.json > /dev/null 2> $O/r2c38_b2b_se76.err; echo "se rc=$?"
grep "^\[b2b\]" $O/r2c38_b2b_se76.err
