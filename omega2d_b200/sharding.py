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
