"""GPU: o2d_advect_bodies -- the reference's convection step with a body (BEM right-hand side, the caller's solve,
find_vels, move, clear_inner_layer per stage; particles resident on the device) against the reference-generated
golden vectors of tests/golden/make_golden_bodies.py."""
import os
import sys

import numpy as np
import pytest

import omega2d_b200 as o2d
from omega2d_b200 import workloads as W

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_bodies import synthetic_solve  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gb():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_bodies.npz"))


@pytest.mark.parametrize("core", ["v2", "rm"])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_advect_bodies_golden(core, order, ctx, gb):
    g = gb
    calls = []
    inner = synthetic_solve(g["ttx"], g["tty"], g["tnx"], g["tny"])

    def solve(stage, time, pu, pv):
        calls.append((stage, time))
        return inner(stage, time, pu, pv)
    z = np.zeros(143, np.float32)
    ips, dt, fs = float(g["ips"][0]), float(g["dt"][0]), tuple(g["fs"])
    x, y, u, v, vs, ss, fx, fy, fu, fv = ctx.advect_bodies_np(core, order, g["x"], g["y"], g["s"], g["r"],
                                                              (g["nx"], g["ny"], g["idx"], z, z), g["area"], (g["tnx"], g["tny"]),
                                                              solve, dt, ips, fs, fldpt=(g["fx"], g["fy"]))
    want = g[f"{core}_o{order}"]
    # one solve per stage, at the reference's stage times
    frac = {1: [0.0], 2: [0.0, 2.0 / 3.0], 3: [0.0, 0.5, 0.75]}[order]
    assert [c[0] for c in calls] == list(range(order))
    np.testing.assert_allclose([c[1] for c in calls], [f * dt for f in frac], rtol=1e-12)
    # the solved strengths of the last stage and the stage velocities: plain tolerance
    assert W.rel_err((vs, ss), tuple(g[f"{core}_o{order}_str"])) < 1e-5
    assert W.rel_err((u, v), (want[2], want[3])) < 1e-5
    # positions: the push of clear_inner_layer is discontinuous at its threshold (a particle that sits within rounding of
    # "0.001 ips below the layer" is moved by one implementation and left by the other); everything else agrees to ~1 ulp
    d = np.hypot(x - want[0], y - want[1])
    flips = d > 2e-6
    assert flips.mean() <= 2e-3, (flips.sum(), d.max())
    assert d.max() <= 0.0011 * ips + 2e-6, d.max()
    # and the step did push particles out: nothing is left inside the clear layer
    apo = 0.5 * np.cos(np.pi / 143)
    cutoff = 0.5 / np.sqrt(2.0 * np.pi)
    assert (np.hypot(x, y) > apo + (cutoff - 0.0011) * ips - 1e-6).all()
    assert (np.hypot(g["x"], g["y"]) < apo + (cutoff - 0.0011) * ips).sum() > 20     # the input had particles in it
    # the inert field points that rode along: same checks against the reference's tracers
    wf = g[f"{core}_o{order}_fld"]
    assert W.rel_err((fu, fv), (wf[2], wf[3])) < 1e-5
    df = np.hypot(fx - wf[0], fy - wf[1])
    assert (df > 2e-6).mean() <= 5e-3 and df.max() <= 0.0011 * ips + 2e-6, (df.max(), (df > 2e-6).sum())
    assert (np.hypot(fx, fy) > apo + (cutoff - 0.0011) * ips - 1e-6).all()
    # ... and without them the particles come out bit-identical
    alone = ctx.advect_bodies_np(core, order, g["x"], g["y"], g["s"], g["r"], (g["nx"], g["ny"], g["idx"], z, z), g["area"],
                                 (g["tnx"], g["tny"]), inner, dt, ips, fs)
    assert np.array_equal(alone[0], x) and np.array_equal(alone[1], y) and np.array_equal(alone[2], u)


def test_advect_bodies_argument_errors(ctx, gb):
    g = gb
    z = np.zeros(143, np.float32)
    solve = synthetic_solve(g["ttx"], g["tty"], g["tnx"], g["tny"])
    with pytest.raises(o2d.O2DError) as ei:
        ctx.advect_bodies_np("v2", 4, g["x"], g["y"], g["s"], g["r"], (g["nx"], g["ny"], g["idx"], z, z), g["area"],
                             (g["tnx"], g["tny"]), solve, 0.01, 0.02)
    assert ei.value.code == o2d.ERR_INVALID

    def failing(stage, time, pu, pv):
        raise RuntimeError("solver diverged")
    with pytest.raises(o2d.O2DError):
        ctx.advect_bodies_np("v2", 2, g["x"], g["y"], g["s"], g["r"], (g["nx"], g["ny"], g["idx"], z, z), g["area"],
                             (g["tnx"], g["tny"]), failing, 0.01, 0.02)
