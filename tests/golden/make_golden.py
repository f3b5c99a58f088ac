"""Generate tests/golden/ref_golden.npz from the UNMODIFIED reference build (oracle/_ref).

Run in the build container, where /root/reference exists:
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these are outputs of the
reference's own templates (points_affect_points<float,double> etc., reference
src/Influence.h, src/InfluenceVort.h) on small seeded inputs, stored together with the
inputs so the fixtures do not depend on numpy's RNG stream.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from omega2d_b200 import workloads as W  # noqa: E402


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    out = {}
    # particles: clustered blob + a few coincident/self pairs, per-particle radii
    sx, sy, ss, sr = W.gaussian_blob(700, seed=11, signed=True)
    ss = (ss * 700).astype(np.float32)
    sr = (sr * 3).astype(np.float32)
    tx, ty, _, tr = W.gaussian_blob(257, seed=12)
    tr = (tr * 3).astype(np.float32)
    # first 40 targets coincide with sources (self pairs: zero velocity, finite vorticity)
    tx[:40], ty[:40], tr[:40] = sx[:40], sy[:40], sr[:40]
    out.update(sx=sx, sy=sy, ss=ss, sr=sr, tx=tx, ty=ty, tr=tr)
    # targets away from sources for the EXP gradient variants (NaN at d=0)
    out.update(txo=tx[40:].copy(), tyo=ty[40:].copy(), tro=tr[40:].copy())
    for core in oracle.CORES:
        R = oracle.ref(core)
        for blob in (0, 1):
            trr = tr if blob else None
            u, v = R.points_on_points(oracle.VELONLY, sx, sy, ss, sr, tx, ty, trr)
            out[f"{core}_u_{blob}"] = np.stack([u, v])
            tq_tx, tq_ty, tq_tr = (tx, ty, trr) if core != "exp" else (tx[40:], ty[40:], None if trr is None else tr[40:])
            u, v, w = R.points_on_points(oracle.VELANDVORT, sx, sy, ss, sr, tq_tx, tq_ty, tq_tr)
            out[f"{core}_uw_{blob}"] = np.stack([u, v, w])
            u, v, e = R.points_on_points(oracle.VELANDSHEAR, sx, sy, ss, sr, tq_tx, tq_ty, tq_tr)
            out[f"{core}_ue_{blob}"] = np.stack([u, v, e])
        tq_tx, tq_ty, tq_tr = (tx, ty, tr) if core != "exp" else (tx[40:], ty[40:], tr[40:])
        out[f"{core}_ug_1"] = np.stack(R.points_on_points_grad(sx, sy, ss, sr, tq_tx, tq_ty, tq_tr))
        out[f"{core}_flops"] = np.array([R.flops(i) for i in range(9)], np.int32)
    R = oracle.ref("v2")
    # panels: the C2 circle (143 panels) + field points and blobs around it
    nx, ny, idx, vs, sps = W.circle_panels(143)
    px, py, ps, pr = W.annulus_blobs(400, seed=21)
    out.update(nx=nx, ny=ny, idx=idx, vs=vs, sps=sps, px=px, py=py, ps=ps, pr=pr)
    out["pan_vs"] = np.stack(R.panels_on_points(nx, ny, idx, vs, sps, px, py))
    out["pan_v"] = np.stack(R.panels_on_points(nx, ny, idx, vs, None, px, py))
    pu, pv, area = R.points_on_panels(px, py, ps, nx, ny, idx)
    out["ptpan"] = np.stack([pu, pv])
    out["area"] = area
    # Gaussian vorticity
    out["vort"] = R.points_vorticity(sx, sy, ss, sr, tx, ty)
    # accumulate semantics: non-zero initial velocities
    u0 = np.linspace(-1, 1, tx.size).astype(np.float32)
    v0 = np.linspace(2, -2, tx.size).astype(np.float32)
    out["u0"], out["v0"] = u0, v0
    out["v2_u_1_accum"] = np.stack(R.points_on_points(oracle.VELONLY, sx, sy, ss, sr, tx, ty, tr, u0, v0))
    # finalize
    fu, fv, fw, _ = R.finalize_vels(out["v2_uw_1"][0], out["v2_uw_1"][1], out["v2_uw_1"][2], None, (1.0, -0.5))
    out["fin"] = np.stack([fu, fv, fw])
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
