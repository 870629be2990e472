#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3"
BWAGPU_SORT_JOBS=1 $K --tag sort2M > $O/r2c31_sort.json 2> $O/r2c31_sort.err; echo "sort rc=$?"
