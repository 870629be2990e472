"""CPU fuzz of k_search's LOGIC alone (tests/host_emu, one emulated thread at a time) against the live reference: random option
sets (zero penalties included), read shapes on both sides of 255 bases, repeat-rich reads (long hit lists), small private arenas
(overflow into the pooled passes) and few slots.   python scripts/fuzz_thread_logic.py <seed> <iterations>
TEST INFRASTRUCTURE."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo')
import refload as R
abi = R.abi
golden = np.load('/root/repo/tests/golden/aln_golden.npz')
T = golden["genome"]
idx = R.bwa.index.build_index(T)
ridx = R.RefIndex(idx)
he = R.emu().emu_load_index(ridx.arr)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 30
t0 = time.time()
for it in range(n_it):
    zero = int(rng.integers(0, 6))  # 0..2: that penalty is zero
    kw = dict(
        s_mm=0 if zero == 0 else int(rng.integers(1, 6)), s_gapo=0 if zero == 1 else int(rng.integers(1, 14)),
        s_gape=0 if zero == 2 else int(rng.integers(1, 7)),
        max_gapo=int(rng.integers(0, 3)), max_gape=int(rng.integers(0, 8)),
        indel_end_skip=int(rng.integers(0, 7)), max_del_occ=int(rng.integers(1, 20)),
        seed_len=int(rng.choice([8, 16, 32, 1024])), max_seed_diff=int(rng.integers(0, 3)),
        max_top2=int(rng.choice([0, 1, 3, 30])), max_entries=int(rng.choice([200, 5000, 2000000])),
        mode=int(rng.choice([0x01, 0x00, 0x05, 0x11, 0x15, 0x04, 0x10])) | 0x02,
    )
    if rng.random() < 0.5:
        kw["fnr"] = float(rng.choice([0.04, 0.01, 0.1]))
    else:
        kw["fnr"] = -1.0; kw["max_diff"] = int(rng.integers(0, 5))
    opt = abi.default_gap_opt(**kw)
    shape = int(rng.integers(0, 4))
    if shape == 0:
        lo = int(rng.integers(8, 60)); hi = lo + int(rng.integers(0, 60))
    elif shape == 1:
        lo, hi = 240, 290   # both sides of 255 bases
    else:
        lo = int(rng.integers(20, 40)); hi = lo + 10
    reads = R.bwa.simulate.simulate_reads(T, 40, (lo, hi), seed=int(rng.integers(1, 1 << 30)), sub_rate=float(rng.choice([0.0, 0.02, 0.06])),
                                          n_rate=float(rng.choice([0.0, 0.01, 0.05])))
    want = R.ref_aln(ridx, reads, opt, threads=4)
    cap1 = int(rng.choice([32, 128, 1024, 4096])); slots = int(rng.choice([1, 3, 7])); chunks = int(rng.choice([4, 64]))
    try:
        got = R.emu_aln(he, reads, opt, cap1=cap1, aln_cap1=int(rng.choice([8, 64])), n_slots=slots, pool_chunks=chunks)
    except Exception as e:
        print(it, "EXC", e, kw, cap1, slots, chunks); bad += 1; continue
    errs = R.compare_aln(want, got, "fuzz")
    if errs:
        bad += 1
        print(it, "MISMATCH", kw, (lo, hi), cap1, slots, chunks, errs[:2], flush=True)
print(f"{n_it} iterations, {bad} bad, {time.time() - t0:.0f} s, hits {int(np.sum(want[0]))} in the last")
