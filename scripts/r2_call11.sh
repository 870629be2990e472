#!/bin/bash
# round 2, GPU call 11: ncu -- the launch list of a bench step, full captures of k_search (C4), k_bgzf_deflate, k_global_warp, k_sa
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
: > $O/r2c11_box.log
python -m pytest tests/test_bgzf.py -m gpu -x -q > $O/r2c11_bgzf_test.log 2>&1; echo "bgzf test rc=$?" >> $O/r2c11_box.log
python scripts/bgzf_bench.py 400000 > $O/r2c11_bgzf_bench.json 2> $O/r2c11_bgzf_bench.err; echo "bgzf bench rc=$?" >> $O/r2c11_box.log
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-aln-only --no-parity"
$B > $O/r2c11_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c11_launches.csv $B > $O/r2c11_ncu_launch.log 2>&1
echo "launch list rc=$?" >> $O/r2c11_box.log
$B > $O/r2c11_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_global_warp|k_sa|k_sw" -s 6 -c 3 -o $O/r2c11_prof_small $B > $O/r2c11_ncu_small.log 2>&1
echo "small kernels rc=$?" >> $O/r2c11_box.log
K="python scripts/kbench.py --genome-bp 3100000000 --read-len 100 --reads 2000000 --reps 1"
$K > $O/r2c11_kplain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_search" -s 1 -c 1 -o $O/r2c11_prof_search $K > $O/r2c11_ncu_search.log 2>&1
echo "k_search rc=$?" >> $O/r2c11_box.log
Z="python scripts/bgzf_bench.py 200000"
$Z > $O/r2c11_zplain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_bgzf_deflate" -s 1 -c 1 -o $O/r2c11_prof_bgzf $Z > $O/r2c11_ncu_bgzf.log 2>&1
echo "bgzf ncu rc=$?" >> $O/r2c11_box.log
cat $O/r2c11_box.log; cat $O/r2c11_bgzf_bench.json; tail -2 $O/r2c11_bgzf_test.log
