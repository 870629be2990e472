"""Host-side profile of a bam2bam run WITHOUT a GPU: the batch shim over tests/cpu_stub (the reference's per-record CPU
functions answer the bwa_gpu_* calls), on a small genome.  Only the shim's own `host CPU seconds by activity` line is of
interest here (the stub's alignment time is not): it is what a change to the host stages is measured with before it goes
to the GPU box.  Usage: python scripts/host_profile.py [pairs] [workdir]"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import bamio, refload as R
import test_batched_bam2bam as TB

args = [a for a in sys.argv[1:] if not a.startswith("--")]
pairs = int(args[0]) if args else 100000
wd = args[1] if len(args) > 1 else "/tmp/hp"
os.makedirs(wd, exist_ok=True)
fa = os.path.join(wd, "g.fa")
T = R.bwa.simulate.make_genome(4000000, seed=8, repeat_frac=0.01, max_copies=5)
if not os.path.exists(fa + ".bwt"):
    R.bwa.simulate.write_fasta(fa, T, 3)
    subprocess.run([R.REF_BWA, "index", "-a", "is", fa], check=True, capture_output=True)
bam = os.path.join(wd, f"pe{pairs}.bam")
if not os.path.exists(bam):
    r1, r2 = R.bwa.simulate.simulate_pairs(T, pairs, 100, seed=5, bad_mate_frac=0.05, bad_mate_sub=0.12)
    bamio.write_unaligned_bam_fast(bam, r1, r2)
TB.build_stub()
env = {"BWAGPU_SHIM_THREADS": os.environ.get("BWAGPU_SHIM_THREADS", "8")}
env.update({k: v for k, v in os.environ.items() if k.startswith("BWAGPU_")})
t0 = time.time()
log = TB.run_bam2bam(fa, bam, os.path.join(wd, "out_stub.bam"), "stub", env_extra=env)
print("wall %.1f s" % (time.time() - t0))
for l in log.splitlines():
    if "host CPU seconds" in l or "pipelined stages" in l:
        print(l)
if "--check" in sys.argv:
    ref_out = os.path.join(wd, f"out_cpu{pairs}.bam")
    if not os.path.exists(ref_out):
        TB.run_bam2bam(fa, bam, ref_out, None)
    TB.compare(ref_out, os.path.join(wd, "out_stub.bam"))
    print("identical to the unmodified reference: %d pairs" % pairs)
