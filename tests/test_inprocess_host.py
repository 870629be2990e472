"""`bam2bam` run IN-PROCESS through libbwa_gpu_batch.so -- the way bench.py drives the full pipeline: dlopen the library,
call the reference's own bwa_bam_to_bam through it, several runs per process with the index kept loaded -- and the two
refusals of the batched drop-in: no device (there is no CPU path) and `.sai` side inputs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import bamio
import refload as R

ROOT = R.ROOT
HOSTDIR = os.path.join(ROOT, "integration", "_host")
DRIVER = os.path.join(HOSTDIR, "bwa_host")
SHIM = os.path.join(ROOT, "integration", "libbwa_gpu_batch.so")
STUBDIR = os.path.join(ROOT, "tests", "cpu_stub")
STUB = os.path.join(STUBDIR, "libbwagpu_cpu_stub.so")
INPROC = os.path.join(STUBDIR, "inproc_host")


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    if not (os.path.exists(DRIVER) and os.path.exists(SHIM) and os.path.exists(R.REF_BWA)):
        pytest.skip("integration/_host, oracle/_ref or the batch shim is not built")
    d = tmp_path_factory.mktemp("inproc")
    T = R.bwa.simulate.make_genome(300000, seed=21, repeat_frac=0.05)
    fa = str(d / "g.fa")
    R.bwa.simulate.write_fasta(fa, T, 2)
    subprocess.run([R.REF_BWA, "index", "-a", "is", fa], check=True, capture_output=True)
    r1, r2 = R.bwa.simulate.simulate_pairs(T, 1500, 100, seed=6)
    bam = str(d / "pe.bam")
    bamio.write_unaligned_bam_fast(bam, r1, r2)
    ref_out = str(d / "ref.bam")
    r = subprocess.run([R.REF_BWA, "bam2bam", "-g", fa, "-t", "1", "-f", ref_out, bam], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return d, fa, bam, ref_out


def same_records(a, b):
    x, y = bamio.read_bam_records(a), bamio.read_bam_records(b)
    assert len(x) == len(y) and len(x) > 0
    bad = [i for i, (p, q) in enumerate(zip(x, y)) if p != q]
    assert not bad, (len(bad), bamio.describe(x[bad[0]]), bamio.describe(y[bad[0]]))


INPROC_PY = r"""
import ctypes as C, sys
H = C.CDLL(sys.argv[1])
H.bwa_bam_to_bam.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p]
H.bwa_gpu_batch_keep_index(1)
for run in range(int(sys.argv[5])):
    args = [b"bam2bam", b"-g", sys.argv[2].encode(), b"-t", b"1", b"-f", (sys.argv[4] + f".{run}.bam").encode(), sys.argv[3].encode()]
    av = (C.c_char_p * (len(args) + 1))(*args, None)
    rc = H.bwa_bam_to_bam(len(args), av, b"inproc-test")
    assert rc == 0, rc
    print("run", run, "ok", flush=True)
H.bwa_gpu_batch_drop_index()
"""


def test_inprocess_without_a_device_fails_loudly(work):
    """No GPU here: the in-process run must get as far as the shim's own pass 1 (so dlopen, the entry wrapper, RTLD_NEXT and the
    interposition of sequential_loop_pass1 all work without LD_PRELOAD) and then DIE on bwa_gpu_init -- never align on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, fa, bam, _ = work
    r = subprocess.run([sys.executable, "-c", INPROC_PY, SHIM, fa, bam, str(d / "nodev"), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert "[bwa_gpu_batch] bwa_gpu_init:" in r.stderr and "no CUDA device" in r.stderr, r.stderr[-2000:]
    assert "run 0 ok" not in r.stdout


def test_sai_inputs_are_refused(work):
    d, fa, bam, _ = work
    sai = str(d / "x.sai")
    open(sai, "wb").write(bytes(R.abi.default_gap_opt()))
    env = dict(os.environ, LD_PRELOAD=SHIM)
    r = subprocess.run([DRIVER, "bam2bam", "-g", fa, "-t", "1", "-1", sai, "-f", str(d / "sai.bam"), bam], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode != 0
    assert ".sai inputs" in r.stderr and "not supported" in r.stderr


def test_inprocess_repeated_runs_on_the_stub(work):
    """Three runs in one process with the index kept: every run's BAM equals the plain reference's; only the first loads the index."""
    if not os.path.exists("/root/reference/bwtaln.h") and not os.path.exists(INPROC):
        pytest.skip("inproc_host not built (needs the reference headers once)")
    from test_batched_bam2bam import build_stub
    if not build_stub():
        pytest.skip("cpu stub not built")
    src = os.path.join(STUBDIR, "inproc_host.c")
    if not os.path.exists(INPROC) or os.path.getmtime(INPROC) < max(os.path.getmtime(src), os.path.getmtime(SHIM), os.path.getmtime(STUB)):
        subprocess.run(["gcc", "-O2", "-o", INPROC, src, "-I", os.path.join(ROOT, "integration"), "-Wl,--no-as-needed", STUB, SHIM,
                        "-Wl,-rpath," + STUBDIR, "-Wl,-rpath," + os.path.join(ROOT, "integration"), "-Wl,-rpath," + HOSTDIR,
                        "-Wl,-rpath," + os.path.join(ROOT, "network-aware-bwa_b200"), "-Wl,-rpath-link," + HOSTDIR,
                        "-Wl,-rpath-link," + os.path.join(ROOT, "network-aware-bwa_b200")], check=True)
    d, fa, bam, ref_out = work
    r = subprocess.run([INPROC, "3", "1", fa, bam, str(d / "stub")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("run ")]
    assert len(lines) == 3
    for run in range(3):
        same_records(ref_out, str(d / f"stub.{run}.bam"))
        assert int(lines[run][lines[run].index("reads_aln") + 1]) == 3000
    load = [float(l[l.index("index_load") + 1]) for l in lines]
    assert r.stderr.count("[cpu_stub] bwa_gpu_*") == 1, "the device context must be set up once, not per run"
    assert load[1] < 0.5 * max(load[0], 1e-3) + 0.01 and load[2] < 0.5 * max(load[0], 1e-3) + 0.01


def test_batch_reader_mixed_input_on_the_stub(work):
    """The shim's batch reader (integration/shim_io.c fastin_read_pairs, in place of one read_bam_pair per record): single
    reads, mates in both orders, alignment tags that must leave the record (erase_unwanted_tags, bwaseqio.c:411-464) next to
    tags that stay (RG, a B array), and BGZF blocks so small that records straddle them -- the BAM must equal the plain
    reference's."""
    import struct
    if not os.path.exists(INPROC):
        pytest.skip("inproc_host not built")
    d, fa, _, _ = work
    rng = np.random.default_rng(17)
    genome = R.bwa.simulate.make_genome(300000, seed=21, repeat_frac=0.05)

    def rec(name, seq, flag, tags):
        l = int(seq.size)
        codes = bamio.NT16[seq]
        if l & 1:
            codes = np.append(codes, 0)
        packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
        body = struct.pack("<iiIIiiii", -1, -1, (4680 << 16) | (len(name) + 1), flag << 16, l, -1, -1, 0)
        body += name + b"\0" + packed + bytes([30]) * l + tags
        return struct.pack("<i", len(body)) + body

    junk = (b"NMC\x01" + b"MDZ50A49\0" + b"XTAU" + b"X0C\x01" + b"X1S\x02\x00" + b"YQI\x05\x00\x00\x00" + b"AMC\x25" + b"XAZchr1,+5,100M,0;\0"
            + b"SMc\x25")
    keep = b"RGZrg1\0" + b"BCZACGT\0" + b"ZBBS\x03\x00\x00\x00\x01\x00\x02\x00\x03\x00" + b"xyi\x07\x00\x00\x00"
    out = []
    for i in range(900):
        pos = int(rng.integers(0, genome.size - 400))
        a = genome[pos:pos + 70].copy()
        b = (3 - genome[pos + 200:pos + 270][::-1]).astype(genome.dtype)
        name = f"q{i}".encode()
        tags = keep[: [0, 7, 15, len(keep)][i % 4]] + (junk if i % 3 == 0 else b"") + (keep if i % 5 == 0 else b"")
        kind = i % 4
        if kind == 0:
            out.append(rec(name, a, 4, tags))
        elif kind == 1:
            out += [rec(name, a, 77, tags), rec(name, b, 141, keep)]
        elif kind == 2:
            out += [rec(name, b, 141, tags), rec(name, a, 77, junk + keep)]  # mates in reverse order
        else:
            out += [rec(name, a, 77 | 512, b""), rec(name, b, 141, tags)]      # one mate failed QC
    text = b"@HD\tVN:1.0\tSO:unsorted\n@RG\tID:rg1\tSM:s\n"
    stream = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 0) + b"".join(out)
    bam = str(d / "mixed.bam")
    with open(bam, "wb") as f:
        for o in range(0, len(stream), 701):
            f.write(bamio._bgzf_block(stream[o:o + 701], 1))
        f.write(bamio.BGZF_EOF)
    ref_out = str(d / "mixed.ref.bam")
    r = subprocess.run([R.REF_BWA, "bam2bam", "-g", fa, "-t", "1", "-f", ref_out, bam], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([INPROC, "1", "0", fa, bam, str(d / "mixed.stub")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, BWAGPU_BATCH_RECORDS="200"))
    assert r.returncode == 0, r.stderr[-3000:]
    same_records(ref_out, str(d / "mixed.stub.0.bam"))
    recs = bamio.read_bam_records(ref_out)
    assert len(recs) == 900 + 675 and any(b"ZBB" in x for x in recs) and not any(b"XTAU" in x and b"q3\0" in x for x in recs[:5])


@pytest.mark.gpu
def test_inprocess_gpu_matches_reference(work):
    d, fa, bam, ref_out = work
    r = subprocess.run([sys.executable, "-c", INPROC_PY, SHIM, fa, bam, str(d / "gpu"), "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "run 1 ok" in r.stdout
    for run in range(2):
        same_records(ref_out, str(d / f"gpu.{run}.bam"))
