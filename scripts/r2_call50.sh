#!/bin/bash
# round 2, GPU call 50: full ncu capture of the request-lean k_search on C4 (2 M reads); resident blocks per SM on that kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 3"
for b in 4 5 6; do
  BWAGPU_T1_BLOCKS_PER_SM=$b $K --tag bps$b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['full']
print('%-8s search %.1f ms  width %.1f ms' % (d['tag'], f['ms_search'], f['ms_width']))"
done > $O/r2c50_bps.log 2>&1
cat $O/r2c50_bps.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search -c 1 -o $O/r2c50_prof_search -f \
  python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 1 > $O/r2c50_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/r2c50_prof_search.ncu-rep
