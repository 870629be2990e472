#!/bin/bash
# round 2, GPU call 12: ncu full capture of k_search (thread per read, pass 0) on C4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 1"
$K > $O/r2c12_kplain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_search$' -s 1 -c 1 -o $O/r2c12_prof_search $K > $O/r2c12_ncu_search.log 2>&1
echo "k_search rc=$?" > $O/r2c12_box.log
cat $O/r2c12_box.log; tail -3 $O/r2c12_ncu_search.log
