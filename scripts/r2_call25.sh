#!/bin/bash
# round 2, GPU call 25: BGZF deflate kernel with 512 threads / positions per batch against 256
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
python scripts/bgzf_bench.py 400000 > $O/r2c25_t256.json 2> $O/r2c25_t256.err; echo "256 rc=$?"
BWAGPU_LIB=$PWD/network-aware-bwa_b200/variants/libbwagpu_bgzf512.so python scripts/bgzf_bench.py 400000 > $O/r2c25_t512.json 2> $O/r2c25_t512.err; echo "512 rc=$?"
cat $O/r2c25_t256.json $O/r2c25_t512.json
