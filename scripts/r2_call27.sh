#!/bin/bash
# round 2, GPU call 27 (8 GPUs): bench.py under torchrun as the driver's scaling run launches it
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L | wc -l; nproc; free -g | head -2; df -h /tmp | tail -1; } > $O/r2c27_box.log 2>&1
S=$SECONDS
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 2 --no-aln-only --no-parity > $O/r2c27_bench_8gpu.json 2> $O/r2c27_bench_8gpu.err
echo "bench8 rc=$? wall $((SECONDS-S)) s" >> $O/r2c27_box.log
cp /tmp/bench_host_rank0.log $O/r2c27_bench_host_rank0.log 2>/dev/null
cat $O/r2c27_box.log; tail -4 $O/r2c27_bench_8gpu.err
