"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, align their shards (with the
oracle standing in for the device, which is absent here), and rank 0 reassembles the batch in read
order -- compared with the committed golden vectors."""
import os
import subprocess
import sys

import numpy as np
import pytest

import refload as R

HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r"""
import os, sys
sys.path.insert(0, {tests!r})
import numpy as np, torch.distributed as dist
import refload as R
from test_kernel_logic import golden_case
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=int(sys.argv[2]))
golden = np.load(os.path.join({tests!r}, "golden", "aln_golden.npz"))
idx = R.bwa.index.build_index(golden["genome"])
reads, opt, want = golden_case(golden, "se36")
mine = R.bwa.shard.shard_reads(reads, dist.get_rank(), dist.get_world_size())
local = R.orc_aln(R.orc_index(idx), mine, opt)
full = R.bwa.shard.gather_alignments(local, dist)
if dist.get_rank() == 0:
    errs = R.compare_aln(want, full, "gloo")
    assert not errs, errs
    print("GATHER_OK", full[0].size)
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
"""


def test_shard_bounds_cover_everything():
    sb = R.bwa.shard.shard_bounds
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            parts = [sb(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_matches_golden(tmp_path):
    port = 29500 + os.getpid() % 1000
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(tests=HERE, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "GATHER_OK 1500" in outs[0][0]
