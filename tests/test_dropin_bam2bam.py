"""End-to-end drop-in check with the reference as the HOST: the unmodified `bam2bam` workflow
(bam2bam.c, BAM in -> BAM out, -t 1 = the deterministic oracle mode) runs twice -- once as is, once
with integration/libbwa_gpu_interpose.so pre-loaded, which forwards bwa_cal_sa_reg_gap and bwt_sa to
libbwagpu.so and checks bwa_gpu_mate_sw against every aln_local_core job the run produces.  Every BAM
record (flags, pos, MAPQ, CIGAR, mate fields, XT/NM/X0/X1/XM/XO/XG/MD/XA tags) must be identical."""
import os
import subprocess

import pytest

import bamio
import refload as R

pytestmark = pytest.mark.gpu
ROOT = R.ROOT
DRIVER = os.path.join(R.ROOT, "integration", "_host", "bwa_host")  # the product host build of the unmodified reference
SHIM = os.path.join(ROOT, "integration", "libbwa_gpu_interpose.so")


def run_bam2bam(prefix, bam_in, bam_out, preload, extra=()):
    env = dict(os.environ)
    env["BWAGPU_LANES"] = "1"
    if preload:
        env["LD_PRELOAD"] = SHIM
    r = subprocess.run([DRIVER, "bam2bam", "-g", prefix, "-t", "1", *extra, "-f", bam_out, bam_in], capture_output=True,
                       text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stderr


@pytest.fixture(scope="module")
def genome(tmp_path_factory):
    if not (os.path.exists(DRIVER) and os.path.exists(SHIM) and os.path.exists(R.REF_BWA)):
        pytest.skip("oracle/_ref or the interposer is not built")
    d = tmp_path_factory.mktemp("dropin")
    T = R.bwa.simulate.make_genome(400000, seed=8, repeat_frac=0.08, max_copies=5)
    fa = str(d / "g.fa")
    R.bwa.simulate.write_fasta(fa, T, 3)
    subprocess.run([R.REF_BWA, "index", "-a", "is", fa], check=True, capture_output=True)
    return d, fa, T


def compare(a_path, b_path):
    a, b = bamio.read_bam_records(a_path), bamio.read_bam_records(b_path)
    assert len(a) == len(b) and len(a) > 0
    bad = [i for i, (x, y) in enumerate(zip(a, b)) if x != y]
    assert not bad, (len(bad), bamio.describe(a[bad[0]]), bamio.describe(b[bad[0]]))
    return a


def test_single_end_bam_identical(genome):
    d, fa, T = genome
    reads = R.bwa.simulate.simulate_reads(T, 2500, (36, 76), seed=3, n_rate=0.002)
    bam = str(d / "se.bam")
    bamio.write_unaligned_bam(bam, reads)
    run_bam2bam(fa, bam, str(d / "se_cpu.bam"), preload=False)
    log = run_bam2bam(fa, bam, str(d / "se_gpu.bam"), preload=True)
    assert "[bwa_gpu_interpose] calls: cal_sa_reg_gap=2500" in log
    recs = compare(str(d / "se_cpu.bam"), str(d / "se_gpu.bam"))
    assert sum(b"XA" in r for r in recs) > 0  # repeats were exercised


def test_paired_end_bam_identical_and_sw_checked(genome):
    d, fa, T = genome
    r1, r2 = R.bwa.simulate.simulate_pairs(T, 2000, 70, seed=5, bad_mate_frac=0.15, bad_mate_sub=0.12)
    bam = str(d / "pe.bam")
    bamio.write_unaligned_bam(bam, r1, r2)
    run_bam2bam(fa, bam, str(d / "pe_cpu.bam"), preload=False)
    log = run_bam2bam(fa, bam, str(d / "pe_gpu.bam"), preload=True)
    compare(str(d / "pe_cpu.bam"), str(d / "pe_gpu.bam"))
    line = [l for l in log.splitlines() if l.startswith("[bwa_gpu_interpose] calls:")][-1]
    fields = dict(kv.split("=") for kv in line.replace(";", "").split() if "=" in kv)
    assert int(fields["cal_sa_reg_gap"]) == 4000 and int(fields["bwt_sa"]) > 4000
    assert int(fields["checked"]) > 20 and int(fields["mismatches"]) == 0  # K5 vs aln_local_core on the run's own jobs


def test_adna_options_bam_identical(genome):
    """BASELINE.json config 5 options: seeding off, -n 0.01 -o 2."""
    d, fa, T = genome
    reads = R.bwa.simulate.simulate_reads(T, 1200, (30, 50), seed=9, adna=True, sub_rate=0.01)
    bam = str(d / "adna.bam")
    bamio.write_unaligned_bam(bam, reads)
    extra = ("-l", "1024", "-n", "0.01", "-o", "2")
    run_bam2bam(fa, bam, str(d / "adna_cpu.bam"), preload=False, extra=extra)
    run_bam2bam(fa, bam, str(d / "adna_gpu.bam"), preload=True, extra=extra)
    compare(str(d / "adna_cpu.bam"), str(d / "adna_gpu.bam"))
