#!/bin/bash
# round 2, GPU call 26: ncu full captures of k_bgzf_inflate_warp and k_global_warp
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
Z="python scripts/inflate_bench.py 400000"
$Z > $O/r2c26_inflate.json 2> $O/r2c26_inflate.err &&
ncu --set full --clock-control none --import-source on -k regex:"k_bgzf_inflate_warp" -s 1 -c 1 -o $O/r2c26_prof_inflate $Z > $O/r2c26_ncu_inflate.log 2>&1
echo "inflate ncu rc=$?"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-aln-only --no-parity"
$B > $O/r2c26_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_global_warp" -s 8 -c 1 -o $O/r2c26_prof_gwarp $B > $O/r2c26_ncu_gwarp.log 2>&1
echo "gwarp ncu rc=$?"
cat $O/r2c26_inflate.json
