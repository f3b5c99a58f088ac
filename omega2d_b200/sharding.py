"""Contiguous-block sharding of particles over ranks.

Rank p owns particles [p*B, min((p+1)*B, N)) with B = ceil(N / P): its slice of the sources
AND of the targets/outputs.  Sources are replicated to every rank once per evaluation (one
all-gather of the packed block, 16 B per particle); outputs never leave their rank.
"""
from __future__ import annotations


def block_size(n: int, nranks: int) -> int:
    """Per-rank block capacity B (the `ns_block` argument of o2d_points_on_points_sharded)."""
    return max(1, -(-n // nranks))


def block_range(n: int, nranks: int, rank: int):
    """[begin, end) of rank's contiguous block; empty for trailing ranks when n < nranks*B."""
    b = block_size(n, nranks)
    lo = min(n, rank * b)
    hi = min(n, (rank + 1) * b)
    return lo, hi


def weak_scaling_count(n1: int, nranks: int) -> int:
    """Particle count that keeps pair interactions PER GPU constant: N_P = N_1 * sqrt(P),
    rounded to a multiple of nranks."""
    n = int(round(n1 * nranks ** 0.5))
    return max(nranks, (n // nranks) * nranks)


def gather_layout(ns_block: int, nranks: int, tile: int = 1024):
    """Layout of the all-gathered packed source array (o2d_pp_sharded_device, csrc/o2d_api.cu): rank p's block occupies
    entries [p*blk, (p+1)*blk) with blk = ns_block rounded up to a source PAIR (the kernels read pairs); the array is
    padded to whole `tile`-source tiles behind the last block.  Returns (blk, padded_total)."""
    blk = (ns_block + 1) // 2 * 2
    total = -(-(blk * nranks) // tile) * tile
    return blk, total


#: a rank without particles fills its block with zero-strength entries far from any flow and with a small radius
FILLER = (1.0e6, 1.0e6, 0.0, 1.0e-3)


def pad_block(arrays, count: int):
    """What pack_sources_kernel writes for entries beyond the rank's last particle: zero-strength copies of that
    particle (they add exactly 0 wherever its own term is finite), or FILLER when the rank owns none.
    arrays = (x, y, s, r) of the rank's particles; returns four arrays of length `count`."""
    import numpy as np
    k = len(arrays[0])
    out = []
    for row, a in enumerate(arrays):
        b = np.empty(count, np.float32)
        b[:k] = a
        b[k:] = (a[k - 1] if row != 2 else 0.0) if k else FILLER[row]
        out.append(b)
    return out
