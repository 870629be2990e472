#!/bin/bash
# round 2, GPU call 35: whole bam2bam runs of the round-1 shapes with the round-2 shim (SE 76 bp / 100 Mb genome)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python scripts/bam2bam_bench.py --mode se --reads 4000000 --len 76 --genome-bp 100000000 --cpu-sample 100000 --out $O/r2c35_b2b_se76.json > /dev/null 2> $O/r2c35_b2b_se76.err; echo "se rc=$?"
cat $O/r2c35_b2b_se76.json | cut -c1-1500
tail -5 $O/r2c35_b2b_se76.err
