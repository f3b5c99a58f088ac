"""Generate tests/golden/ref_golden_substep.npz from the UNMODIFIED reference build (oracle/_ref).

    python tests/golden/make_golden_substep.py        (in the build container: needs /root/reference)

Outputs of the reference's own containers and methods -- InfluenceVisitor, zero_vels,
finalize_vels, Points::move -- driven by the find_vels / advect orchestration restated in
oracle/ref_driver.cpp (Convection.h itself needs Eigen), stored with their inputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from omega2d_b200 import workloads as W  # noqa: E402

FS = (1.0, -0.25)
DT = 0.01


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    out = {}
    x, y, s, r = W.c1_speedtest(900, seed=77)            # uniform radius, strengths U[-1,1]
    r2 = (r * np.random.default_rng(5).uniform(0.8, 1.2, r.size)).astype(np.float32)   # per-particle radii
    nx, ny, idx, pvs, pss = W.circle_panels(48, radius=0.5, seed=78)
    tx, ty, _, tr = W.onbody_random(211, seed=79)
    out.update(x=x, y=y, s=s, r=r, r2=r2, nx=nx, ny=ny, idx=idx, pvs=pvs, pss=pss, tx=tx, ty=ty, tr=tr,
               fs=np.array(FS), dt=np.array(DT))
    for core in oracle.CORES:
        R = oracle.ref(core)
        for res, key in ((oracle.VELONLY, "u"), (oracle.VELANDVORT, "uw"), (oracle.VELANDSHEAR, "ue")):
            out[f"{core}_fv_{key}_blob"] = np.stack(R.find_vels(res, (x, y, s, r2), (nx, ny, idx, pvs, pss), tx, ty, tr, FS))
            out[f"{core}_fv_{key}_inert"] = np.stack(R.find_vels(res, (x, y, s, r2), (nx, ny, idx, pvs, pss), tx, ty, None, FS))
        out[f"{core}_fv_u_nopanels"] = np.stack(R.find_vels(oracle.VELONLY, (x, y, s, r2), None, tx, ty, tr, FS))
        for order in (1, 2, 3):
            out[f"{core}_adv{order}"] = np.stack(R.advect(order, x, y, s, r, DT, None, FS))
            out[f"{core}_adv{order}_r2"] = np.stack(R.advect(order, x, y, s, r2, DT, None, FS))
        out[f"{core}_adv2_panels"] = np.stack(R.advect(2, x, y, s, r2, DT, (nx, ny, idx, pvs, pss), FS))
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_substep.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
