#!/bin/bash
# round 2, GPU call 36: why is a whole SE run slower per read than a PE one?  Full logs + stage trace of the SE shape, variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python scripts/bam2bam_bench.py --mode se --reads 4000000 --len 76 --genome-bp 100000000 --cpu-sample 20000 \
  --log-dir $O/r2c36_logs --gpu-env trace:BWAGPU_TRACE=1 --gpu-env hostinflate:BWAGPU_HOST_INFLATE=1 --gpu-env batch256k:BWAGPU_BATCH_RECORDS=262144 \
  --gpu-env nomallopt:BWAGPU_MALLOPT=0 --out $O/r2c36_b2b_se76.json > /dev/null 2> $O/r2c36_b2b_se76.err; echo "se rc=$?"
grep "^\[b2b\]" $O/r2c36_b2b_se76.err
