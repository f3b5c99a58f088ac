"""Generate tests/golden/ref_golden_reflect.npz from the UNMODIFIED reference build (oracle/_ref): outputs of the
reference's own clear_inner_layer<float> (method 1, src/Reflect.h:490-768) and reflect_interior<float> (:174-380)
on its own Collections.

    python tests/golden/make_golden_reflect.py        (in the build container: needs /root/reference)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from omega2d_b200 import workloads as W  # noqa: E402

CF = 0.5 / math.sqrt(2.0 * math.pi)     # the factor Convection passes (src/Convection.h:306)
IPS = 0.02191                           # flow_over_circle's inter-particle spacing


def square_body(m=40, half=0.5):
    """closed square, clockwise like the reference's bodies, m panels per side: has corners, where a particle's
    closest element is a node shared by two panels (exact distance ties)"""
    c = np.array([[-half, -half], [-half, half], [half, half], [half, -half]], np.float32)
    pts = np.concatenate([c[k] + (c[(k + 1) % 4] - c[k]) * (np.arange(m, dtype=np.float32)[:, None] / m) for k in range(4)])
    n = pts.shape[0]
    idx = np.empty(2 * n, np.uint32); idx[0::2] = np.arange(n); idx[1::2] = (np.arange(n) + 1) % n
    return np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1]), idx


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    R = oracle.ref("v2")
    rng = np.random.default_rng(31)
    out = {"cf": np.float32(CF), "ips": np.float32(IPS)}
    n = 6000
    # circle: particles inside, in the layer, and outside
    cx, cy, cidx, _, _ = W.circle_panels(143, radius=0.5)
    rho = np.sqrt(rng.uniform(0.0, 0.62 ** 2, n)); th = rng.uniform(0, 2 * math.pi, n)
    px = (rho * np.cos(th)).astype(np.float32); py = (rho * np.sin(th)).astype(np.float32)
    s = np.ones(n, np.float32); r = np.full(n, 0.043818, np.float32)
    x, y, _ = R.clear_inner_layer(cx, cy, cidx, px, py, CF, IPS, s, r)
    out.update(circ_nx=cx, circ_ny=cy, circ_idx=cidx, circ_px=px, circ_py=py, circ_out=np.stack([x, y]))
    x, y = R.clear_inner_layer(cx, cy, cidx, px, py, CF, IPS)          # inert field points
    out["circ_out_fldpt"] = np.stack([x, y])
    # square with corners
    sx, sy, sidx = square_body()
    qx = rng.uniform(-0.65, 0.65, n).astype(np.float32); qy = rng.uniform(-0.65, 0.65, n).astype(np.float32)
    x, y, _ = R.clear_inner_layer(sx, sy, sidx, qx, qy, CF, IPS, s, r)
    out.update(sq_nx=sx, sq_ny=sy, sq_idx=sidx, sq_px=qx, sq_py=qy, sq_out=np.stack([x, y]))
    # reflect_interior: same bodies; particles inside, near and outside.  The square set is seeded with the
    # exact-tie cases: the centre (four nodes, eight hits, normals cancel -> hits[0] fallback, :305-317), points
    # on the axes and diagonals (two-sided ties), a point exactly on a node and one exactly on a panel.
    x, y = R.reflect_interior(cx, cy, cidx, px, py, s, r)
    out["circ_reflect"] = np.stack([x, y])
    special = np.array([[0.0, 0.0], [0.0, 0.25], [0.25, 0.0], [-0.125, 0.0], [0.25, 0.25], [-0.3, 0.3], [0.45, 0.45],
                        [0.5, 0.5], [0.5, 0.0125], [0.4875, -0.5], [0.0, -0.4], [0.49, 0.0]], np.float32)
    rx = np.concatenate([special[:, 0], qx[special.shape[0]:]]); ry = np.concatenate([special[:, 1], qy[special.shape[0]:]])
    x, y = R.reflect_interior(sx, sy, sidx, rx, ry, s, r)
    out.update(sq_rx=rx, sq_ry=ry, sq_reflect=np.stack([x, y]))
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_reflect.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
