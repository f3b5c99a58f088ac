"""Seeded synthetic particle / panel sets for the BASELINE.json configs.

Inputs are generated with numpy so that the CUDA path, the C oracle and the compiled
reference all see bit-identical float32 arrays (the reference's own scene generators seed
from std::random_device, reference src/FlowFeature.cpp:738-751, so whole-run comparison
is meaningless -- SURVEY F9).  Nothing here computes velocities.
"""
from __future__ import annotations

import math

import numpy as np


def c1_speedtest(n: int = 20000, seed: int = 12345):
    """C1, examples/speedtest20kx10.json restated: N blobs uniform in [-2,2]^2, strengths
    U[-1,1], uniform radius vdelta = 2*sqrt(6)*sqrt(0.005/100) (reference
    src/Simulation.cpp:83-85,889; BlockOfRandom scatter src/FlowFeature.cpp:775)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2.0, 2.0, n).astype(np.float32)
    y = rng.uniform(-2.0, 2.0, n).astype(np.float32)
    s = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    r = np.full(n, 2.0 * math.sqrt(6.0) * math.sqrt(0.005 / 100.0), np.float32)
    return x, y, s, r


def gaussian_blob(n: int, seed: int = 2024, sigma: float = 0.25, vary_radius: bool = True,
                  signed: bool = False):
    """C3/C4/C5 particle set: positions ~ N(0, sigma^2) isotropic, strength 1/N (or
    U[-1,1]/N), radius 1.5/sqrt(N) optionally x U[0.8,1.2] (per-particle radii)."""
    rng = np.random.default_rng(seed)
    x = rng.normal(0.0, sigma, n).astype(np.float32)
    y = rng.normal(0.0, sigma, n).astype(np.float32)
    if signed:
        s = (rng.uniform(-1.0, 1.0, n) / n).astype(np.float32)
    else:
        s = np.full(n, 1.0 / n, np.float32)
    r = np.full(n, 1.5 / math.sqrt(n))
    if vary_radius:
        r = r * rng.uniform(0.8, 1.2, n)
    return x, y, s, r.astype(np.float32)


def onbody_random(n: int, seed: int = 7):
    """The onbody driver's recipe (reference extern/onbody/src/main2dvort.cpp:72-80):
    x,y ~ U[-1,1], s ~ U[-1,1]/N, r = (0.6+U[0,1))/sqrt(N)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    y = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    s = (rng.uniform(-1.0, 1.0, n) / n).astype(np.float32)
    r = ((0.6 + rng.uniform(0.0, 1.0, n)) / math.sqrt(n)).astype(np.float32)
    return x, y, s, r


def circle_panels(npan: int, radius: float = 0.5, center=(0.0, 0.0), seed: int = 99,
                  clockwise: bool = True):
    """One closed circle of `npan` panels; node order clockwise like the reference's
    SolidCircle (reference src/BoundaryFeature.cpp:489-525).  Returns node x,y (npan
    nodes), idx (2*npan uint32), vortex strengths vs and source strengths ss ~ U[-1,1]."""
    rng = np.random.default_rng(seed)
    k = np.arange(npan)
    th = 2.0 * math.pi * k / npan
    if clockwise:
        th = -th
    nx = (center[0] + radius * np.cos(th)).astype(np.float32)
    ny = (center[1] + radius * np.sin(th)).astype(np.float32)
    idx = np.empty(2 * npan, np.uint32)
    idx[0::2] = k
    idx[1::2] = (k + 1) % npan
    vs = rng.uniform(-1.0, 1.0, npan).astype(np.float32)
    ss = rng.uniform(-1.0, 1.0, npan).astype(np.float32)
    return nx, ny, idx, vs, ss


def multi_circle_panels(ncircles: int, npan_each: int, radius: float = 0.1, spacing: float = 1.2,
                        seed: int = 99):
    """C4 panel set: `ncircles` closed circles of `npan_each` panels, centres on a line."""
    xs, ys, idxs, vss, sss = [], [], [], [], []
    off = 0
    for c in range(ncircles):
        cx = (c - 0.5 * (ncircles - 1)) * spacing
        nx, ny, idx, vs, ss = circle_panels(npan_each, radius, (cx, 1.5), seed + c)
        xs.append(nx); ys.append(ny); idxs.append(idx + np.uint32(off)); vss.append(vs); sss.append(ss)
        off += nx.size
    return (np.concatenate(xs), np.concatenate(ys), np.concatenate(idxs).astype(np.uint32),
            np.concatenate(vss), np.concatenate(sss))


def annulus_blobs(n: int, rin: float = 0.52, rout: float = 1.5, seed: int = 5,
                  radius: float = 0.043818):
    """C2 stand-in: blobs in an annulus around the unit-diameter circle with the
    flow_over_circle core radius vdelta."""
    rng = np.random.default_rng(seed)
    rho = np.sqrt(rng.uniform(rin * rin, rout * rout, n))
    th = rng.uniform(0.0, 2.0 * math.pi, n)
    x = (rho * np.cos(th)).astype(np.float32)
    y = (rho * np.sin(th)).astype(np.float32)
    s = (rng.uniform(-1.0, 1.0, n) * 1e-3).astype(np.float32)
    r = np.full(n, radius, np.float32)
    return x, y, s, r


def strided_subset(n: int, count: int):
    """Indices of `count` targets strided through n (onbody's ntskip method, reference
    extern/onbody/src/main2dvort.cpp:116-133)."""
    count = min(count, n)
    stride = max(1, n // count)
    return np.arange(0, stride * count, stride, dtype=np.int64)[:count]


def rel_err(got, want):
    """north_star tolerance metric: max_i ||d_i|| / max_i ||want_i|| over vector (or scalar)
    fields given as tuples of component arrays."""
    if not isinstance(got, (tuple, list)):
        got, want = (got,), (want,)
    g = np.stack([np.asarray(a, np.float64) for a in got])
    w = np.stack([np.asarray(a, np.float64) for a in want])
    num = np.sqrt(((g - w) ** 2).sum(axis=0)).max() if g.shape[1] else 0.0
    den = np.sqrt((w ** 2).sum(axis=0)).max() if w.shape[1] else 1.0
    if not np.isfinite(num):
        return float("inf")
    return float(num / den) if den > 0 else float(num)
