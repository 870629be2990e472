#!/bin/bash
# round 2, GPU call 39: (a) stage + lane trace of two bench steps (where do pass 1 / pass 2 wait?), (b) ncu full capture of k_sw with the DPX cell
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f /tmp/bench_host_rank0.log
BWAGPU_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c39_bench_trace.json 2> $O/r2c39_bench_trace.err; echo "bench rc=$?"
cp /tmp/bench_host_rank0.log $O/r2c39_bench_trace_host.log 2>/dev/null
Z="python scripts/sw_bench.py"
ncu --set full --clock-control none --import-source on -k regex:"k_sw" -s 1 -c 1 -o $O/r2c39_prof_sw $Z > $O/r2c39_ncu_sw.log 2>&1
echo "sw ncu rc=$?"
ls -la $O/r2c39_prof_sw.ncu-rep
