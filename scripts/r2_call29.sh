#!/bin/bash
# round 2, GPU call 29: smoke(), then the bench as the driver runs it (K = 20, W = 5), both arms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c29_smoke.log 2>&1; echo "smoke rc=$?" > $O/r2c29_box.log
S=$SECONDS
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2c29_bench_ref.json 2> $O/r2c29_bench_ref.err
echo "ref rc=$? wall $((SECONDS-S)) s" >> $O/r2c29_box.log
S=$SECONDS
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2c29_bench.json 2> $O/r2c29_bench.err
echo "bench rc=$? wall $((SECONDS-S)) s" >> $O/r2c29_box.log
cp /tmp/bench_host_rank0.log $O/r2c29_bench_host.log 2>/dev/null
tail -2 $O/r2c29_smoke.log; cat $O/r2c29_box.log
