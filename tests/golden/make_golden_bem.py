"""Generate tests/golden/ref_golden_bem.npz from the UNMODIFIED reference build (oracle/_ref):
outputs of the reference's own panels_on_panels_coeff<float> (src/Coefficients.h:155-573) and
vels_to_rhs_panels<float> (src/RHS.h:54-116) on small seeded surfaces, stored with their inputs.

    python tests/golden/make_golden_bem.py        (in the build container: needs /root/reference)
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
    R = oracle.ref("v2")     # the coefficient path does not depend on the core function
    out = {}
    # A: the flow_over_circle body (143 panels, unit diameter); B: a smaller off-centre circle; C: an ellipse-ish
    # body with unequal panel lengths (node spacing modulated), counter-clockwise
    ax, ay, aidx, _, _ = W.circle_panels(143, radius=0.5)
    bx, by, bidx, _, _ = W.circle_panels(61, radius=0.2, center=(1.1, 0.3))
    k = np.arange(97)
    th = 2 * np.pi * (k / 97.0 + 0.05 * np.sin(2 * np.pi * k / 97.0))
    cx = (-1.4 + 0.6 * np.cos(th)).astype(np.float32); cy = (0.25 * np.sin(th)).astype(np.float32)
    cidx = np.empty(2 * 97, np.uint32); cidx[0::2] = k; cidx[1::2] = (k + 1) % 97
    out.update(ax=ax, ay=ay, aidx=aidx, bx=bx, by=by, bidx=bidx, cx=cx, cy=cy, cidx=cidx)
    surf = {"a": (ax, ay, aidx), "b": (bx, by, bidx), "c": (cx, cy, cidx)}
    for s in "abc":
        c, bases, sa = R.panels_on_panels_coeff(surf[s])
        out[f"coeff_{s}{s}"] = c
        out[f"bases_{s}"] = np.stack(bases)
        for t in "abc":
            if t != s:
                out[f"coeff_{s}{t}"] = R.panels_on_panels_coeff(surf[s], surf[t])[0]
    for s, t in (("a", "a"), ("a", "c"), ("c", "b")):
        bases = out[f"bases_{t}"]; sa = out[f"bases_{s}"][4]
        out[f"coeffv_{s}{t}"] = R.panels_on_panels_coeff_vortex(surf[s], surf[t], bases, sa, s == t)
    rng = np.random.default_rng(17)
    pu = rng.normal(size=143).astype(np.float32); pv = rng.normal(size=143).astype(np.float32)
    nbc = (0.1 * rng.normal(size=143)).astype(np.float32)
    out.update(pu=pu, pv=pv, nbc=nbc, rhs_a=R.vels_to_rhs_panels(ax, ay, aidx, pu, pv, nbc))
    # Surfaces::add_rot_strengths_base on a body-bound copy of C (off-centre body, unequal panels), rate 0.7
    vs0 = rng.normal(size=97).astype(np.float32); ss0 = rng.normal(size=97).astype(np.float32)
    rvs, rss, ctr = R.add_rot_strengths(cx, cy, cidx, 0.7, vs0, ss0)
    out.update(rot_vs0=vs0, rot_ss0=ss0, rot_vs=rvs, rot_ss=rss, rot_center=np.array(ctr, np.float32), rot_rate=np.float32(0.7))
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_bem.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
