#!/bin/bash
# round 2, GPU call 18: device inflate -- parity tests, shim tests, bench with it and with the zlib workers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_bgzf.py -m gpu -x -q > $O/r2c18_bgzf.log 2>&1; echo "bgzf rc=$?" > $O/r2c18_box.log
timeout 900 python -m pytest tests/test_inprocess_host.py tests/test_batched_bam2bam.py tests/test_dropin_bam2bam.py tests/test_gpu_parity.py -m gpu -x -q -k "inprocess or gpu_ or chunking or bam_identical" > $O/r2c18_host.log 2>&1; echo "host rc=$?" >> $O/r2c18_box.log
timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only > $O/r2c18_bench.json 2> $O/r2c18_bench.err
echo "bench rc=$?" >> $O/r2c18_box.log
cp /tmp/bench_host_rank0.log $O/r2c18_bench_host.log 2>/dev/null
BWAGPU_HOST_INFLATE=1 timeout 1200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-aln-only --no-parity > $O/r2c18_bench_hostinf.json 2> $O/r2c18_bench_hostinf.err
echo "bench hostinf rc=$?" >> $O/r2c18_box.log
tail -3 $O/r2c18_bgzf.log; tail -3 $O/r2c18_host.log
grep -E "pipelined|host CPU" $O/r2c18_bench_host.log | head -21 | tail -3
cat $O/r2c18_box.log
