#!/bin/bash
# round 2, GPU call 2: the new batched shim + in-process bench on C4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; nproc; free -g | head -2; } > $O/r2c2_box.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c2_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c2_box.log
timeout 1500 python bench.py --steps 3 --warmup 2 > $O/r2c2_bench.json 2> $O/r2c2_bench.err
echo "bench rc=$?" >> $O/r2c2_box.log
cp /tmp/bench_host_rank0.log $O/r2c2_bench_host.log 2>/dev/null
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c2_bench_ref.json 2> $O/r2c2_bench_ref.err
echo "ref rc=$?" >> $O/r2c2_box.log
timeout 900 python bench.py --steps 3 --warmup 2 --pairs 2000000 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c2_bench_2M.json 2> $O/r2c2_bench_2M.err
echo "bench2M rc=$?" >> $O/r2c2_box.log
cp /tmp/bench_host_rank0.log $O/r2c2_bench_host_2M.log 2>/dev/null
tail -n 5 $O/r2c2_pytest.log; cat $O/r2c2_box.log; tail -n 3 $O/r2c2_bench.err
