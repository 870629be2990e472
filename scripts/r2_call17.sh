#!/bin/bash
# round 2, GPU call 17: k_search on C4 with cache-policy variants (index loads streaming, context loads kept)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3 --check 10000"
for v in base idx1 idx2 ctx idx1ctx idx2ctx; do
  if [ "$v" = "base" ]; then unset BWAGPU_LIB; else export BWAGPU_LIB=$PWD/network-aware-bwa_b200/variants/libbwagpu_$v.so; fi
  $K --tag $v > $O/r2c17_$v.json 2> $O/r2c17_$v.err; echo "$v rc=$?"
done
