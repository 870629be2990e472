#!/bin/bash
# round 2, GPU call 37: whole bam2bam runs of the round-1 shapes with the round-2 shim, BGZF input (SE 76 bp / 100 Mb genome; aDNA options)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python scripts/bam2bam_bench.py --mode se --reads 4000000 --len 76 --genome-bp 100000000 --cpu-sample 100000 \
  --log-dir $O/r2c37_logs --gpu-env again:BWAGPU_X=1 --out $O/r2c37_b2b_se76.json > /dev/null 2> $O/r2c37_b2b_se76.err; echo "se rc=$?"
grep "^\[b2b\]" $O/r2c37_b2b_se76.err
