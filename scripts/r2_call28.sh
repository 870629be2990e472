#!/bin/bash
# round 2, GPU call 28 (8 GPUs): bench.py under torchrun after the blocking-sync change
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L | wc -l; nproc; } > $O/r2c28_box.log 2>&1
S=$SECONDS
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 3 --warmup 2 --no-aln-only --no-parity > $O/r2c28_bench_8gpu.json 2> $O/r2c28_bench_8gpu.err
echo "bench8 rc=$? wall $((SECONDS-S)) s" >> $O/r2c28_box.log
S=$SECONDS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/r2c28_bench_8gpu_ref.json 2> $O/r2c28_bench_8gpu_ref.err
echo "ref8 rc=$? wall $((SECONDS-S)) s" >> $O/r2c28_box.log
cat $O/r2c28_box.log
