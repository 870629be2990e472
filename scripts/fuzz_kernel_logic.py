"""CPU fuzz of the kernel LOGIC (tests/host_emu: k_search one thread at a time, k_search_warp on a coroutine warp) against the live
reference over random option sets:   python scripts/fuzz_kernel_logic.py <seed> <iterations>"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import refload as R
abi = R.abi
golden = np.load('/root/repo/tests/golden/aln_golden.npz')
T = golden["genome"]
idx = R.bwa.index.build_index(T)
ridx = R.RefIndex(idx)
h = R.wemu().wemu_load_index(ridx.arr)
h4 = R.wemu(4).wemu_load_index(ridx.arr)
he = R.emu().emu_load_index(ridx.arr)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    kw = dict(
        s_mm=int(rng.integers(1, 6)), s_gapo=int(rng.integers(1, 14)), s_gape=int(rng.integers(1, 7)),
        max_gapo=int(rng.integers(0, 3)), max_gape=int(rng.integers(0, 8)),
        indel_end_skip=int(rng.integers(0, 7)), max_del_occ=int(rng.integers(1, 20)),
        seed_len=int(rng.choice([8, 16, 32, 1024])), max_seed_diff=int(rng.integers(0, 3)),
        max_top2=int(rng.choice([0, 1, 3, 30])), max_entries=int(rng.choice([200, 5000, 2000000])),
        mode=int(rng.choice([0x01, 0x00, 0x05, 0x11, 0x15, 0x04, 0x10])) | 0x02,
    )
    if rng.random() < 0.5:
        kw["fnr"] = float(rng.choice([0.04, 0.01, 0.1, 0.001]))
    else:
        kw["fnr"] = -1.0; kw["max_diff"] = int(rng.integers(0, 5))
    opt = abi.default_gap_opt(**kw)
    lo = int(rng.integers(8, 60)); hi = lo + int(rng.integers(0, 60))
    reads = R.bwa.simulate.simulate_reads(T, 50, (lo, hi), seed=int(rng.integers(1, 1 << 30)), sub_rate=float(rng.choice([0.0, 0.02, 0.06])),
                                          n_rate=float(rng.choice([0.0, 0.01, 0.05])))
    want = R.ref_aln(ridx, reads, opt, threads=4)
    t = time.time()
    try:
        got = R.wemu_aln(h, reads, opt)
    except Exception as e:
        print(it, "EXC", e, kw); bad += 1; continue
    errs = R.compare_aln(want, got, "fuzz") if got[4] == 0 else [f"dry {got[4]}"]
    got4 = R.wemu_aln(h4, reads, opt, team=4)
    errs += R.compare_aln(want, got4, "fuzz-team") if got4[4] == 0 else [f"team dry {got4[4]}"]
    try:
        got2 = R.emu_aln(he, reads, opt)
        errs2 = R.compare_aln(want, got2, "fuzz-thread")
    except Exception as e:
        errs2 = ["EXC " + str(e)[:80]]
    if errs or errs2:
        bad += 1
        print(it, "MISMATCH warp:", errs[:1], "thread:", errs2[:1], kw, (lo, hi), flush=True)
    else:
        print(it, "ok", round(time.time() - t, 1), "s max_me", int(want[1].max()), flush=True)
print("bad", bad)
