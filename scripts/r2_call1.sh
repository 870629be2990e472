#!/bin/bash
# round 2, GPU call 1: native index builder (tests + 3.1 Gb timing), C4 parity, aln-only bench on the 3.1 Gb genome
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{ nvidia-smi -L; nproc; free -g; df -h /tmp /dev/shm; } > $O/r2c1_box.log 2>&1
timeout 900 python -m pytest tests/test_index_gpu.py -m gpu -x -q -k "not c4" > $O/r2c1_index_tests.log 2>&1
echo "index tests rc=$?" >> $O/r2c1_box.log
BWAGPU_TRACE=1 timeout 900 python - > $O/r2c1_index31.log 2>&1 <<'PY'
import importlib, time
bwa = importlib.import_module("network-aware-bwa_b200")
t = time.time()
p = bwa.workload.ensure_genome_files(3_100_000_000, 1, 0)
print("ensure 3.1e9:", time.time() - t, p)
t = time.time()
p = bwa.workload.ensure_genome_files(3_050_000_000, 1, 0)
print("ensure 3.05e9:", time.time() - t, p)
PY
echo "index31 rc=$?" >> $O/r2c1_box.log
timeout 900 python -m pytest tests/test_index_gpu.py -m gpu -x -q -k "c4" > $O/r2c1_c4_tests.log 2>&1
echo "c4 tests rc=$?" >> $O/r2c1_box.log
timeout 900 python bench.py --steps 3 --warmup 2 --reads 4000000 --genome-bp 3100000000 --read-len 100 --no-extras > $O/r2c1_bench_31.json 2> $O/r2c1_bench_31.err
echo "bench31 rc=$?" >> $O/r2c1_box.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_index_gpu.py > $O/r2c1_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2c1_box.log
tail -3 $O/r2c1_index_tests.log $O/r2c1_index31.log $O/r2c1_c4_tests.log $O/r2c1_pytest.log; cat $O/r2c1_box.log
