"""Generate tests/golden/ref_golden_bodies.npz: one convection step of a flow WITH a body (flow_over_circle shape).

Run in the build container, where /root/reference exists:
    python tests/golden/make_golden_bodies.py

Orders 1-3 (advect_1st / advect_2nd_ralston / advect_3rd, reference src/Convection.h:277-595) including the per-stage
solve_bem right-hand side (points_affect_panels + Surfaces::finalize_vels, src/BEMHelper.h:88-100) and
clear_inner_layer<float>(1, ...) after every move, driven by oracle/ref_driver.cpp::o2dref_advect_bodies around the
reference's own templates.  The linear solve (Eigen, absent here) is replaced by the deterministic stand-in
`synthetic_solve` below -- the same callback the CUDA path is given in the tests: strengths that cancel the tangential
and normal velocity the particles and the freestream induce at the panel centres.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from omega2d_b200 import workloads as W  # noqa: E402

IPS = np.float32(0.02191)      # sqrt(6) * sqrt(0.02/250): examples/flow_over_circle.json
DT = 0.02
FS = (1.0, 0.0)


def synthetic_solve(ttx, tty, tnx, tny):
    ttx, tty, tnx, tny = (np.asarray(a, np.float32) for a in (ttx, tty, tnx, tny))

    def solve(stage, time, pu, pv):
        return (-(pu * ttx + pv * tty)).astype(np.float32), (-(pu * tnx + pv * tny)).astype(np.float32)
    return solve


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    out = {}
    nx, ny, idx, _, _ = W.circle_panels(143, radius=0.5)
    ttx, tty, tnx, tny, area = oracle.port().panel_bases(nx, ny, idx)
    # particles from just inside the surface (they get pushed out) to one radius away
    x, y, s, r = W.annulus_blobs(3000, rin=0.495, rout=1.0, seed=17)
    # inert Lagrangian field points (tracers) around the body, some inside the clear layer: they ride along in every stage
    fx, fy, _, _ = W.annulus_blobs(400, rin=0.497, rout=0.8, seed=23)
    out.update(nx=nx, ny=ny, idx=idx, ttx=ttx, tty=tty, tnx=tnx, tny=tny, area=area, x=x, y=y, s=s, r=r, fx=fx, fy=fy,
               ips=np.array([IPS]), dt=np.array([DT]), fs=np.array(FS))
    solve = synthetic_solve(ttx, tty, tnx, tny)
    zeros = np.zeros(143, np.float32)
    for core in ("v2", "rm"):
        R = oracle.ref(core)
        for order in (1, 2, 3):
            res = R.advect_bodies(order, x, y, s, r, (nx, ny, idx, zeros, zeros), solve, DT, float(IPS), FS, time=0.0, fldpt=(fx, fy))
            out[f"{core}_o{order}"] = np.stack(res[:4])
            out[f"{core}_o{order}_str"] = np.stack(res[4:6])
            out[f"{core}_o{order}_fld"] = np.stack(res[6:10])
            # the particles must not depend on whether tracers ride along
            alone = R.advect_bodies(order, x, y, s, r, (nx, ny, idx, zeros, zeros), solve, DT, float(IPS), FS, time=0.0)
            assert all(np.array_equal(a, b) for a, b in zip(alone[:6], res[:6]))
            moved = np.hypot(res[0] - x, res[1] - y)
            print(core, order, "max displacement", moved.max(), "vs", float(np.abs(res[4]).max()))
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_bodies.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
