#!/bin/bash
# round 2, GPU call 8: k_search on pipeline-sized batches -- longest-first job order, pop caps
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py"
C4="--genome-bp 3100000000 --read-len 100 --reads 1048576 --batches 65536,131072,262144,524288"
$K $C4 --tag base > $O/r2c8_base.json 2> $O/r2c8_base.err; echo "base rc=$?" > $O/r2c8_box.log
BWAGPU_SORT_JOBS=1 $K $C4 --check 20000 --tag sort > $O/r2c8_sort.json 2> $O/r2c8_sort.err; echo "sort rc=$?" >> $O/r2c8_box.log
BWAGPU_SORT_JOBS=1 BWAGPU_POP_CAP=4096 $K $C4 --tag sort_cap4096 > $O/r2c8_sort_cap4096.json 2> $O/r2c8_sort_cap4096.err; echo "sortcap rc=$?" >> $O/r2c8_box.log
cat $O/r2c8_box.log
