"""The BATCHED drop-in (integration/bwa_gpu_batch.c): the unmodified reference `bam2bam -t 1` with its two
sequential loop functions replaced by batching ones that make one device call per phase.  The output BAM
must be record-identical to the plain CPU run -- flags, pos, MAPQ, CIGAR, mate fields and every tag -- for
single-end, paired-end (pairing, XA lists, mate rescue) and aDNA-option runs, and with batches small enough
that records, SA-row sub-ranges and the pass-2 position cache all straddle batch boundaries.

Two arms share every line of the driver:
  * `-m gpu`:      the shim calls libbwagpu.so (K2/K3, K4, K5+K6 on the B200);
  * `-m "not gpu"`: tests/cpu_stub answers the same bwa_gpu_* calls with the reference's per-record CPU
                    functions, which checks the shim's HOST logic (batching, record/replay, drand48 and
                    cache order) on a box without a GPU."""
import os
import subprocess

import numpy as np
import pytest

import bamio
import refload as R

ROOT = R.ROOT
DRIVER = os.path.join(R.ROOT, "integration", "_host", "bwa_host")  # the product host build of the unmodified reference
SHIM = os.path.join(ROOT, "integration", "libbwa_gpu_batch.so")
STUB = os.path.join(ROOT, "tests", "cpu_stub", "libbwagpu_cpu_stub.so")


def build_stub():
    src = os.path.join(ROOT, "tests", "cpu_stub", "bwagpu_cpu_stub.c")
    if os.path.exists("/root/reference/bwtaln.h") and (not os.path.exists(STUB) or os.path.getmtime(STUB) < os.path.getmtime(src)):
        subprocess.run(["gcc", "-O2", "-w", "-fgnu89-inline", "-fPIC", "-shared", "-I", "/root/reference", "-I",
                        os.path.join(ROOT, "include"), "-o", STUB, src, "-lz"], check=True)
    return os.path.exists(STUB)


def run_bam2bam(prefix, bam_in, bam_out, mode, extra=(), env_extra=None):
    env = dict(os.environ)
    env["BWAGPU_LANES"] = "1"
    if mode == "gpu":
        env["LD_PRELOAD"] = SHIM
    elif mode == "stub":
        env["LD_PRELOAD"] = STUB + ":" + SHIM
    env.update(env_extra or {})
    r = subprocess.run([DRIVER, "bam2bam", "-g", prefix, "-t", "1", *extra, "-f", bam_out, bam_in], capture_output=True,
                       text=True, env=env, timeout=1800)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stderr


@pytest.fixture(scope="module")
def genome(tmp_path_factory):
    if not (os.path.exists(DRIVER) and os.path.exists(SHIM) and os.path.exists(R.REF_BWA)):
        pytest.skip("integration/_host, oracle/_ref or the batch shim is not built")
    d = tmp_path_factory.mktemp("batched")
    # wide repeat families so that intervals >= 1000 rows (the pass-2 position cache, bam2bam.c:743) occur
    T = R.bwa.simulate.make_genome(600000, seed=8, repeat_frac=0.08, max_copies=5)
    fam = T[1000:1100].copy()
    for c in range(1200):
        T[200000 + c * 100: 200000 + (c + 1) * 100] = fam
    # rescue targets: a 3-substitution copy of the 70-mer at X sits one insert size downstream of Y.  `aln` maps a read
    # of X to X only (after the first hit it keeps hits within one more difference, bwtgap.c:175), so the pair
    # (read at Y, read of X) is discordant and bwa_paired_sw1 finds, accepts and FIXES it via the copy (bwape.c:592-627)
    rng = np.random.default_rng(4)
    fix = []
    for t in range(80):
        X = 20000 + 2000 * t
        Y = 400000 + 2000 * t
        copy = T[X:X + 70].copy()
        for q in rng.choice(70, size=3, replace=False):
            copy[q] = (copy[q] + 1 + rng.integers(0, 3)) & 3
        T[Y + 230:Y + 300] = copy
        fix.append((X, Y))
    fa = str(d / "g.fa")
    R.bwa.simulate.write_fasta(fa, T, 3)
    subprocess.run([R.REF_BWA, "index", "-a", "is", fa], check=True, capture_output=True)
    return d, fa, T, fix


def compare(a_path, b_path):
    a, b = bamio.read_bam_records(a_path), bamio.read_bam_records(b_path)
    assert len(a) == len(b) and len(a) > 0
    bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, (len(bad), bamio.describe(a[bad[0]]), bamio.describe(b[bad[0]]))
    return a


def calls(log):
    line = [l for l in log.splitlines() if l.startswith("[bwa_gpu_batch] device calls:")][-1]
    out = {}
    for part in line.split("  "):
        if "=" in part:
            k, v = part.strip().split("=", 1)
            k = k.split()[-1]
            out[k] = int(v.split()[0])
            out[k + "_units"] = int(v.split("(")[1].split()[0])
    return out


def se_case(genome, mode):
    d, fa, T, _ = genome
    reads = R.bwa.simulate.simulate_reads(T, 3000, (36, 76), seed=3, n_rate=0.002)
    bam = str(d / "se.bam")
    bamio.write_unaligned_bam(bam, reads)
    run_bam2bam(fa, bam, str(d / "se_cpu.bam"), None)
    log = run_bam2bam(fa, bam, str(d / f"se_{mode}.bam"), mode, env_extra={"BWAGPU_BATCH_RECORDS": "700"})
    recs = compare(str(d / "se_cpu.bam"), str(d / f"se_{mode}.bam"))
    c = calls(log)
    assert c["cal_sa_reads_gap"] == 5 and c["cal_sa_reads_gap_units"] == 3000  # ceil(3000 / 700) batches, every read once
    assert c["cal_pac_pos_units"] > 2500
    assert c["global_align_units"] > 50  # the gapped hits' CIGARs came through the batched aln_global_core
    assert sum(b"XA" in r for r in recs) > 0
    assert "not a BGZF file" in log  # a plain gzip stream cannot be inflated in parallel: said so, read ahead on one thread


def pe_case(genome, mode):
    d, fa, T, fix = genome
    r1, r2 = R.bwa.simulate.simulate_pairs(T, 2500, 70, seed=5, bad_mate_frac=0.15, bad_mate_sub=0.12)
    b1, b2 = r1.bases.reshape(-1, 70), r2.bases.reshape(-1, 70)
    for t, (X, Y) in enumerate(fix):  # every 30th pair becomes a rescue target
        b1[30 * t] = T[Y:Y + 70]
        b2[30 * t] = 3 - T[X:X + 70][::-1]
    bam = str(d / "pe.bam")
    bamio.write_unaligned_bam_fast(bam, r1, r2)  # BGZF: the shim inflates its blocks on several threads
    cpu_log = run_bam2bam(fa, bam, str(d / "pe_cpu.bam"), None)
    # small record batches AND small SA-row sub-ranges: cache entries are created in one range and reused in later ones
    log = run_bam2bam(fa, bam, str(d / f"pe_{mode}.bam"), mode, env_extra={"BWAGPU_BATCH_RECORDS": "600", "BWAGPU_BATCH_SA": "3000"})
    compare(str(d / "pe_cpu.bam"), str(d / f"pe_{mode}.bam"))
    c = calls(log)
    assert "not a BGZF file" not in log
    assert c["cal_sa_reads_gap_units"] == 5000
    assert c["mate_sw_path_units"] > 20          # mate rescue ran through the batch call
    assert c["cal_pac_pos_units"] > 5000
    fixed = [l for l in cpu_log.splitlines() if "discordant pairs are fixed" in l][-1]
    assert fixed in log                              # same "[bwa_paired_sw] N out of M ... fixed" line
    assert int(fixed.split()[1]) >= 40               # and the accept-and-rewrite branch of mate rescue really ran
    # the intermediate records partly in memory, partly spilled to the reference's temporary file; three host threads
    log = run_bam2bam(fa, bam, str(d / f"pe_{mode}_spill.bam"), mode,
                      env_extra={"BWAGPU_BATCH_RECORDS": "600", "BWAGPU_MEMTEMP_BYTES": "400000", "BWAGPU_SHIM_THREADS": "3"})
    compare(str(d / "pe_cpu.bam"), str(d / f"pe_{mode}_spill.bam"))
    assert "the rest in the temporary file" in log
    # and everything through the temporary file, one host thread
    run_bam2bam(fa, bam, str(d / f"pe_{mode}_file.bam"), mode, env_extra={"BWAGPU_MEMTEMP_BYTES": "0", "BWAGPU_SHIM_THREADS": "1"})
    compare(str(d / "pe_cpu.bam"), str(d / f"pe_{mode}_file.bam"))


def adna_case(genome, mode):
    d, fa, T, _ = genome
    reads = R.bwa.simulate.simulate_reads(T, 1200, (30, 50), seed=9, adna=True, sub_rate=0.01)
    bam = str(d / "adna.bam")
    bamio.write_unaligned_bam(bam, reads)
    extra = ("-l", "1024", "-n", "0.01", "-o", "2")
    run_bam2bam(fa, bam, str(d / "adna_cpu.bam"), None, extra=extra)
    run_bam2bam(fa, bam, str(d / f"adna_{mode}.bam"), mode, extra=extra)
    compare(str(d / "adna_cpu.bam"), str(d / f"adna_{mode}.bam"))


# ---- the batching 0MQ worker (run_worker_thread replaced): `bam2bam -t 1 -p PORT` keeps the reference's reader, multiplexor
# and output threads and its wire format; only the worker behind inproc://work_io batches.  `remote` = the same worker
# inside a separate `bwa worker` process talking to an untouched `bam2bam -t 0 -p PORT` over TCP.
def free_port():
    import socket
    for base in range(41000, 60000, 7):
        ok = True
        for p in (base, base + 1, base + 2):
            with socket.socket() as s:
                try:
                    s.bind(("127.0.0.1", p))
                except OSError:
                    ok = False
        if ok:
            return base
    raise RuntimeError("no free port triple")


def preload_env(mode, env_extra=None):
    env = dict(os.environ)
    env["BWAGPU_LANES"] = "1"
    env["LD_PRELOAD"] = SHIM if mode == "gpu" else STUB + ":" + SHIM
    env.update(env_extra or {})
    return env


def run_worker_mode(prefix, bam_in, bam_out, mode, remote, env_extra=None):
    port = free_port()
    if not remote:
        r = subprocess.run([DRIVER, "bam2bam", "-g", prefix, "-t", "1", "-p", str(port), "-f", bam_out, bam_in], capture_output=True,
                           text=True, env=preload_env(mode, env_extra), timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
        return r.stderr
    master = subprocess.Popen([DRIVER, "bam2bam", "-g", prefix, "-t", "0", "-p", str(port), "-f", bam_out, bam_in],
                              stderr=subprocess.PIPE, text=True)
    worker = subprocess.Popen([DRIVER, "worker", "-t", "1", "-h", "127.0.0.1", "-p", str(port)], stderr=subprocess.PIPE, text=True,
                              env=preload_env(mode, env_extra))
    try:
        _, merr = master.communicate(timeout=300)
        _, werr = worker.communicate(timeout=120)
    finally:
        for p in (master, worker):
            if p.poll() is None:
                p.kill()
    assert master.returncode == 0, merr[-3000:]
    return werr


def worker_stats(log):
    line = [l for l in log.splitlines() if l.startswith("[run_worker_thread] exiting:")][-1]
    w = line.split()
    return {"batches": int(w[2]), "records": int(w[4]), "dupes": int(w[6])}


def worker_case(genome, mode, remote):
    d, fa, T, fix = genome
    tag = f"{mode}_{'remote' if remote else 'local'}"
    # single-end
    reads = R.bwa.simulate.simulate_reads(T, 3000, (36, 76), seed=3, n_rate=0.002)
    bam = str(d / "w_se.bam")
    bamio.write_unaligned_bam(bam, reads)
    run_bam2bam(fa, bam, str(d / "w_se_cpu.bam"), None)
    log = run_worker_mode(fa, bam, str(d / f"w_se_{tag}.bam"), mode, remote, {"BWAGPU_WORKER_RECORDS": "700"})
    compare(str(d / "w_se_cpu.bam"), str(d / f"w_se_{tag}.bam"))
    st = worker_stats(log)
    assert st["records"] == 2 * 3000 and st["batches"] < st["records"] / 20  # every record once per pass, in batches
    # paired-end with mate rescue and shared position-cache entries
    r1, r2 = R.bwa.simulate.simulate_pairs(T, 2500, 70, seed=5, bad_mate_frac=0.15, bad_mate_sub=0.12)
    b1, b2 = r1.bases.reshape(-1, 70), r2.bases.reshape(-1, 70)
    for t, (X, Y) in enumerate(fix):
        b1[30 * t] = T[Y:Y + 70]
        b2[30 * t] = 3 - T[X:X + 70][::-1]
    bam = str(d / "w_pe.bam")
    bamio.write_unaligned_bam(bam, r1, r2)
    run_bam2bam(fa, bam, str(d / "w_pe_cpu.bam"), None)
    log = run_worker_mode(fa, bam, str(d / f"w_pe_{tag}.bam"), mode, remote, {"BWAGPU_WORKER_RECORDS": "600", "BWAGPU_BATCH_SA": "3000"})
    compare(str(d / "w_pe_cpu.bam"), str(d / f"w_pe_{tag}.bam"))
    st = worker_stats(log)
    assert st["records"] == 2 * 2500 and st["batches"] < st["records"] / 20
    c = calls(log)
    assert c["cal_sa_reads_gap_units"] == 5000 and c["mate_sw_path_units"] > 20


# ---- host logic on the CPU stub (no GPU needed)
@pytest.fixture(scope="module")
def stub():
    if not build_stub():
        pytest.skip("cpu stub not built (needs the reference headers once)")


def test_stub_single_end(genome, stub):
    se_case(genome, "stub")


def test_stub_paired_end(genome, stub):
    pe_case(genome, "stub")


def test_stub_worker_local(genome, stub):
    worker_case(genome, "stub", remote=False)


def test_stub_worker_remote(genome, stub):
    worker_case(genome, "stub", remote=True)


# ---- the real thing
@pytest.mark.gpu
def test_gpu_worker_local(genome):
    worker_case(genome, "gpu", remote=False)


@pytest.mark.gpu
def test_gpu_worker_remote(genome):
    worker_case(genome, "gpu", remote=True)


@pytest.mark.gpu
def test_gpu_single_end(genome):
    se_case(genome, "gpu")


@pytest.mark.gpu
def test_gpu_paired_end(genome):
    pe_case(genome, "gpu")


@pytest.mark.gpu
def test_gpu_adna_options(genome):
    adna_case(genome, "gpu")
