#!/bin/bash
# round 2, GPU call 23: k_search on C4 with more resident blocks per SM (register caps)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3 --check 10000"
for v in base mb8 mb9 mb10; do
  if [ "$v" = "base" ]; then unset BWAGPU_LIB; else export BWAGPU_LIB=$PWD/network-aware-bwa_b200/variants/libbwagpu_$v.so; fi
  $K --tag $v > $O/r2c23_$v.json 2> $O/r2c23_$v.err; echo "$v rc=$?"
done
