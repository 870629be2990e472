"""The C-ABI: struct layouts against the reference's own headers, exported symbols."""
import ctypes as C
import json
import os
import subprocess

import pytest

import refload as R

abi = R.abi
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "abi_layout.json")

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "bwtaln.h"
#include "bwtgap.h"
#define S(t) printf("\"sizeof_" #t "\": %zu,\n", sizeof(t))
#define O(t, f) printf("\"" #t "." #f "\": %zu,\n", offsetof(t, f))
int main(void) {
  printf("{\n");
  S(bwt_t); S(bwt_aln1_t); S(bwt_multi1_t); S(bwa_seq_t); S(gap_opt_t); S(bwt_width_t); S(gap_entry_t);
  O(bwt_t, bwt); O(bwt_t, cnt_table); O(bwt_t, sa_intv); O(bwt_t, n_sa); O(bwt_t, sa);
  O(bwa_seq_t, seq); O(bwa_seq_t, rseq); O(bwa_seq_t, score); O(bwa_seq_t, n_aln); O(bwa_seq_t, aln);
  O(bwa_seq_t, n_multi); O(bwa_seq_t, multi); O(bwa_seq_t, sa); O(bwa_seq_t, pos); O(bwa_seq_t, n_cigar);
  O(bwa_seq_t, cigar); O(bwa_seq_t, tid); O(bwa_seq_t, bc); O(bwa_seq_t, md); O(bwa_seq_t, max_entries);
  O(gap_opt_t, fnr); O(gap_opt_t, max_diff); O(gap_opt_t, seed_len); O(gap_opt_t, max_top2); O(gap_opt_t, trim_qual);
  O(bwt_aln1_t, k); O(bwt_aln1_t, l); O(bwt_aln1_t, score);
  printf("\"end\": 0}\n");
  return 0;
}
"""


def probe_reference(tmp):
    src = os.path.join(tmp, "probe.c")
    open(src, "w").write(PROBE)
    exe = os.path.join(tmp, "probe")
    subprocess.run(["gcc", "-w", "-fgnu89-inline", "-I", "/root/reference", "-o", exe, src], check=True)
    return json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)


def layout_from_ctypes():
    d = {}
    for t in (abi.bwt_t, abi.bwt_aln1_t, abi.bwt_multi1_t, abi.bwa_seq_t, abi.gap_opt_t):
        d[f"sizeof_{t.__name__}"] = C.sizeof(t)
    for t, fields in ((abi.bwt_t, ["bwt", "cnt_table", "sa_intv", "n_sa", "sa"]),
                      (abi.bwa_seq_t, ["seq", "rseq", "score", "n_aln", "aln", "n_multi", "multi", "sa", "pos", "n_cigar",
                                       "cigar", "tid", "bc", "md", "max_entries"]),
                      (abi.gap_opt_t, ["fnr", "max_diff", "seed_len", "max_top2", "trim_qual"]),
                      (abi.bwt_aln1_t, ["k", "l", "score"])):
        for f in fields:
            d[f"{t.__name__}.{f}"] = getattr(t, f).offset
    return d


def test_layout_matches_golden():
    gold = json.load(open(GOLD))
    mine = layout_from_ctypes()
    for k, v in mine.items():
        assert gold[k] == v, k


@pytest.mark.skipif(not os.path.exists("/root/reference/bwtaln.h"), reason="reference headers not present")
def test_golden_matches_reference_headers(tmp_path):
    assert probe_reference(str(tmp_path)) == json.load(open(GOLD))


def test_header_compiles_as_c_and_cxx(tmp_path):
    inc = os.path.join(os.path.dirname(HERE), "include")
    src = tmp_path / "t.c"
    src.write_text('#include "bwa_gpu.h"\nint main(void){ return sizeof(bwa_seq_t) == 200 && sizeof(gap_opt_t) == 64 && '
                   'sizeof(bwt_aln1_t) == 16 ? 0 : 1; }\n')
    for cc, extra in (("gcc", []), ("g++", ["-x", "c++", "-std=c++17"])):
        exe = str(tmp_path / ("t_" + cc))
        subprocess.run([cc, *extra, "-I", inc, "-o", exe, str(src)], check=True)
        assert subprocess.run([exe]).returncode == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/bwtaln.h"), reason="reference headers not present")
def test_header_coexists_with_reference_headers(tmp_path):
    """bwa_gpu.h included after the reference's bwtaln.h must reuse the reference's types."""
    inc = os.path.join(os.path.dirname(HERE), "include")
    src = tmp_path / "t.c"
    src.write_text('#include "bwtaln.h"\n#include "bwa_gpu.h"\nint main(void){ bwa_seq_t s; (void)s; return 0; }\n')
    subprocess.run(["gcc", "-w", "-fgnu89-inline", "-I", "/root/reference", "-I", inc, "-c", "-o",
                    str(tmp_path / "t.o"), str(src)], check=True)


def test_library_exports_every_declared_symbol():
    """No compute calls here (no GPU): just that the .so loads and exports the C-ABI."""
    import re
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "bwa_gpu.h")).read()
    declared = set(re.findall(r"\b(bwa_gpu_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"bwa_gpu_sw_job_t", "bwa_gpu_sw_res_t", "bwa_gpu_stats_t"}
    L = R.bwa.api.lib()
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(R.bwa.api.EXPORTS)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.bwa.api.BwaGpuError):
        R.bwa.api.init()
    opt = abi.default_gap_opt()
    import numpy as np
    with pytest.raises(R.bwa.api.BwaGpuError):
        R.bwa.api.aln_flat(np.zeros(4, np.uint8), np.array([0, 4], np.int64), opt)
