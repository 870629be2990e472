#!/bin/bash
# round 2, GPU call 24: k_search on C4 with FEWER resident blocks per SM
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3"
for b in 7 6 5 4; do
  BWAGPU_T1_BLOCKS_PER_SM=$b $K --tag bps$b > $O/r2c24_bps$b.json 2> $O/r2c24_bps$b.err; echo "bps $b rc=$?"
done
