#!/bin/bash
# round 2, GPU call 10 (2 GPUs): the in-process multi-device path (bwa_gpu_init(n, ids)) and bench.py under torchrun
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; nproc; } > $O/r2c10_box.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" -rs > $O/r2c10_multidev.log 2>&1; echo "multidev rc=$?" >> $O/r2c10_box.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $O/r2c10_bench_2gpu.json 2> $O/r2c10_bench_2gpu.err
echo "bench2 rc=$?" >> $O/r2c10_box.log
# the batched shim on both devices inside one process (BWAGPU_NDEV=2)
BWAGPU_NDEV=2 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c10_bench_ndev2.json 2> $O/r2c10_bench_ndev2.err
echo "ndev2 rc=$?" >> $O/r2c10_box.log
tail -4 $O/r2c10_multidev.log; cat $O/r2c10_box.log; tail -3 $O/r2c10_bench_2gpu.err
