"""Generate tests/golden/ref_golden_large.npz: BASELINE-size parity fixtures from the UNMODIFIED reference build.

Run in the build container, where /root/reference exists (takes a minute or two):
    python tests/golden/make_golden_large.py

What it holds (inputs at these sizes come from the seeded generators in omega2d_b200/workloads.py; the TARGET
coordinates and the indices of the sampled targets are stored, so a change of numpy's RNG stream is detected by the
tests instead of silently moving the goalposts):
  * C4 (BASELINE configs[3]): 4 194 304-particle Gaussian blob, Rosenhead-Moore core + 20 000 panels (4 circles of
    5 000).  For 512 strided targets: the reference's whole find_vels (zero, points_affect_points,
    panels_affect_points, finalize_vels); its panels_affect_points alone; and the float64 evaluation of the same panel
    formulas (oracle/f64.py) -- for RANDOM sheet strengths U[-1,1] (the bench's C4 set) and for SMOOTH BEM-like ones
    (vs = 2 sin(theta) + 0.3, ss = 0.5 cos(theta) on every circle).
    The panel term is where the literal "1e-5 of the reference" rule cannot hold: the reference's float32
    atan2f-atan2f / logf form (src/Kernels.h:262-268,338) is itself 1e-5 ... 3e-3 from exact arithmetic for these short
    panels, which is why both sides are committed (README "Parity rule for panels").
  * C5-like: 2 097 152 particles, Winckelmans-Leonard core, and the same count with the Vatistas-2 core (the
    reference's default): points_affect_points<float,double> velocities on 256 strided targets.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.f64 import panels_on_points_f64  # noqa: E402
from omega2d_b200 import workloads as W  # noqa: E402

N_C4 = 4_194_304
N_2M = 2_097_152


def smooth_strengths(nx, ny, ncircles, npan_each, spacing=1.2):
    """BEM-like sheet strengths: the potential-flow solution shape on every circle."""
    vs = np.empty(ncircles * npan_each, np.float32)
    ss = np.empty_like(vs)
    for c in range(ncircles):
        cx = (c - 0.5 * (ncircles - 1)) * spacing
        sl = slice(c * npan_each, (c + 1) * npan_each)
        th = np.arctan2(ny[sl] - 1.5, nx[sl] - cx)
        vs[sl] = 2.0 * np.sin(th) + 0.3
        ss[sl] = 0.5 * np.cos(th)
    return vs, ss


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    oracle.set_threads()
    out = {}
    # ---- C4 -----------------------------------------------------------------------------------------------
    sx, sy, ss, sr = W.gaussian_blob(N_C4, seed=2024)
    sub = W.strided_subset(N_C4, 512)
    tx, ty, tr = sx[sub].copy(), sy[sub].copy(), sr[sub].copy()
    out.update(c4_sub=sub, c4_tx=tx, c4_ty=ty, c4_tr=tr,
               c4_input_checksum=np.array([sx.astype(np.float64).sum(), sy.astype(np.float64).sum(),
                                           ss.astype(np.float64).sum(), sr.astype(np.float64).sum()]))
    R = oracle.ref("rm")
    nx, ny, idx, vs_r, ss_r = W.multi_circle_panels(4, 5000)
    vs_s, ss_s = smooth_strengths(nx, ny, 4, 5000)
    for tag, pvs, pss in (("random", vs_r, ss_r), ("smooth", vs_s, ss_s)):
        panels = (nx, ny, idx, pvs, pss)
        u, v = R.find_vels(oracle.VELONLY, (sx, sy, ss, sr), panels, tx, ty, tr)
        pu, pv = R.panels_on_points(nx, ny, idx, pvs, pss, tx, ty)
        eu, ev = panels_on_points_f64(nx, ny, idx, pvs, pss, tx, ty)
        out[f"c4_{tag}_vs"] = pvs; out[f"c4_{tag}_ss"] = pss
        out[f"c4_{tag}_find_vels"] = np.stack([u, v])
        out[f"c4_{tag}_panels_ref"] = np.stack([pu, pv])
        out[f"c4_{tag}_panels_f64"] = np.stack([eu, ev])
        print(tag, "reference panel term vs float64:", W.rel_err((pu, pv), (eu, ev)))
    # particles alone (un-normalised sums), to pin the N^2 part separately
    u, v = R.points_on_points(oracle.VELONLY, sx, sy, ss, sr, tx, ty, tr)
    out["c4_particles_ref"] = np.stack([u, v])
    # ---- 2M particles, WL and V2 ---------------------------------------------------------------------------------
    sx, sy, ss, sr = W.gaussian_blob(N_2M, seed=2024)
    sub = W.strided_subset(N_2M, 256)
    out.update(m2_sub=sub, m2_tx=sx[sub].copy(), m2_ty=sy[sub].copy(), m2_tr=sr[sub].copy(),
               m2_input_checksum=np.array([sx.astype(np.float64).sum(), sy.astype(np.float64).sum(),
                                           ss.astype(np.float64).sum(), sr.astype(np.float64).sum()]))
    for core in ("wl", "v2"):
        u, v = oracle.ref(core).points_on_points(oracle.VELONLY, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])
        out[f"m2_{core}"] = np.stack([u, v])
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_large.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
