#!/bin/bash
# round 2, GPU call 13: k_search on C4 vs the L2 fetch-granularity hint
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3"
for g in 32 64 128; do
  BWAGPU_L2_FETCH=$g $K --tag l2fetch$g > $O/r2c13_l2f$g.json 2> $O/r2c13_l2f$g.err; echo "l2fetch $g rc=$?"
done
