"""Read sharding across ranks (one process per GPU) -- SURVEY.md §8(e).

Records are independent and every rank holds a full index replica, so the only multi-rank
logic is: give rank r a contiguous slice of the batch, run the hot path locally, and put the
results back in read order on rank 0 (what the reference's mux does by `recno`,
bam2bam.c:1610-1648).  There is no data-path collective: the gather below moves results only.
Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int) -> tuple:
    """Contiguous, balanced, order-preserving: rank r gets [n*r/world, n*(r+1)/world)."""
    return n * rank // world, n * (rank + 1) // world


def shard_reads(reads, rank: int, world: int):
    """Slice a simulate.Reads batch for this rank."""
    lo, hi = shard_bounds(reads.n, rank, world)
    offs = reads.offs[lo:hi + 1] - reads.offs[lo]
    bases = reads.bases[reads.offs[lo]:reads.offs[hi]]
    return type(reads)(bases, offs, None if reads.pos is None else reads.pos[lo:hi],
                       None if reads.strand is None else reads.strand[lo:hi])


def gather_alignments(local, dist=None, dst: int = 0):
    """local = (n_aln, max_entries, aln_off, aln) of this rank's shard (api.aln_flat's format).
    Returns the whole batch's tuple on rank `dst` (None elsewhere), shards concatenated in rank
    order = read order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    n_aln, max_entries, _, aln = local[:4]
    payload = (np.asarray(n_aln), np.asarray(max_entries), np.asarray(aln))
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst)
    if rank != dst:
        return None
    n_all = np.concatenate([p[0] for p in out])
    m_all = np.concatenate([p[1] for p in out])
    a_all = np.concatenate([p[2] for p in out])
    off = np.zeros(n_all.size + 1, dtype=np.int64)
    off[1:] = np.cumsum(n_all)
    return n_all, m_all, off, a_all
