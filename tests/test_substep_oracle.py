"""CPU: find_vels / advect restatement (oracle/o2d_oracle.c) against the reference-generated golden
vectors (tests/golden/ref_golden_substep.npz) and, where oracle/_ref is built, against the
reference's own containers on fresh inputs."""
import os

import numpy as np
import pytest

import oracle
from omega2d_b200 import workloads as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_substep.npz"))


def _pan(g):
    return (g["nx"], g["ny"], g["idx"], g["pvs"], g["pss"])


@pytest.mark.parametrize("core", oracle.CORES)
def test_find_vels_golden(core, gs, port):
    g = gs
    fs = tuple(g["fs"])
    vort = (g["x"], g["y"], g["s"], g["r2"])
    for res, key in ((oracle.VELONLY, "u"), (oracle.VELANDVORT, "uw"), (oracle.VELANDSHEAR, "ue")):
        for tr, tag in ((g["tr"], "blob"), (None, "inert")):
            got = port.find_vels(core, res, vort, _pan(g), g["tx"], g["ty"], tr, fs)
            want = g[f"{core}_fv_{key}_{tag}"]
            assert W.rel_err(got[:2], (want[0], want[1])) < 2e-6
            if len(got) == 3:
                assert W.rel_err(got[2], want[2]) < 3e-6
    got = port.find_vels(core, oracle.VELONLY, vort, None, g["tx"], g["ty"], g["tr"], fs)
    assert W.rel_err(got, tuple(g[f"{core}_fv_u_nopanels"])) < 2e-6


def check_step(got, want, x0, y0, tol=1e-5):
    """positions: error of the DISPLACEMENT within tol of the largest displacement, plus one float32 ulp of
    the largest coordinate per component (positions are stored in float32: a last-bit difference in a
    velocity can flip the rounding of x + dt*u); velocities: rel_err."""
    dx = np.stack([got[0] - x0, got[1] - y0]).astype(np.float64)
    dw = np.stack([want[0] - x0, want[1] - y0]).astype(np.float64)
    disp = np.sqrt((dw ** 2).sum(axis=0)).max()
    ulp = float(np.spacing(np.float32(max(np.abs(x0).max(), np.abs(y0).max()))))
    perr = np.sqrt(((dx - dw) ** 2).sum(axis=0)).max()
    assert perr < tol * disp + 1.5 * ulp, ("displacement", perr / disp)
    assert W.rel_err((got[2], got[3]), (want[2], want[3])) < tol


@pytest.mark.parametrize("core", oracle.CORES)
def test_advect_golden(core, gs, port):
    g = gs
    fs, dt = tuple(g["fs"]), float(g["dt"])
    for order in (1, 2, 3):
        for rk, tag in (("r", ""), ("r2", "_r2")):
            got = port.advect(core, order, g["x"], g["y"], g["s"], g[rk], dt, None, fs)
            check_step(got, g[f"{core}_adv{order}{tag}"], g["x"], g["y"], 5e-6)
    got = port.advect(core, 2, g["x"], g["y"], g["s"], g["r2"], dt, _pan(g), fs)
    check_step(got, g[f"{core}_adv2_panels"], g["x"], g["y"], 5e-6)


def test_advect_order_of_accuracy(port):
    """Two equal Rosenhead-Moore vortices orbit their midpoint at Omega = Gamma / (pi (d^2 + 2 r^2));
    the error against that analytic orbit shrinks with the integrator's order."""
    import math
    x = np.array([-0.5, 0.5], np.float32); y = np.zeros(2, np.float32)
    s = np.array([1.0, 1.0], np.float32); r = np.full(2, 0.05, np.float32)
    omega = 1.0 / (math.pi * (1.0 + 2 * 0.05 ** 2))
    err = []
    for order in (1, 2, 3):
        xs, ys = x.copy(), y.copy()
        for _ in range(20):
            xs, ys, _, _ = port.advect("rm", order, xs, ys, s, r, 0.5)
        th = omega * 0.5 * 20
        err.append(float(np.hypot(xs[1] - 0.5 * math.cos(th), ys[1] - 0.5 * math.sin(th))))
    assert err[0] > 20 * err[1] and err[1] > 8 * err[2], err


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("core", ["v2", "exp"])
def test_port_matches_reference_fresh_inputs(core, port):
    R = oracle.ref(core)
    x, y, s, r = W.gaussian_blob(800, seed=41, signed=True)
    s = (s * 800).astype(np.float32)
    pan = W.circle_panels(32, radius=0.2, center=(1.5, 0.0), seed=42)
    for order in (1, 2, 3):
        a = R.advect(order, x, y, s, r, 0.005, pan, (0.5, 0.1))
        b = port.advect(core, order, x, y, s, r, 0.005, pan, (0.5, 0.1))
        check_step(b, a, x, y, 5e-6)
