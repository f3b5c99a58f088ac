"""CPU: closed-form sanity checks of the oracle (SURVEY 8c 'analytic sanity checks')."""
import math

import numpy as np
import pytest

import oracle
from omega2d_b200 import workloads as W


def test_single_rm_blob(port):
    # raw (un-normalised) RM velocity of one blob: u = -G dy/(d^2+s^2), v = G dx/(d^2+s^2)
    sx = np.array([0.1], np.float32); sy = np.array([-0.3], np.float32)
    g = np.array([1.7], np.float32); sr = np.array([0.05], np.float32)
    tx = np.array([0.6, -0.2], np.float32); ty = np.array([0.4, 0.1], np.float32)
    u, v = port.points_on_points("rm", oracle.VELONLY, sx, sy, g, sr, tx, ty, None)
    dx, dy = tx - sx, ty - sy
    k = g / (dx * dx + dy * dy + sr * sr)
    np.testing.assert_allclose(u, -k * dy, rtol=1e-6)
    np.testing.assert_allclose(v, k * dx, rtol=1e-6)


@pytest.mark.parametrize("core", oracle.CORES)
def test_gradient_identities(core, port):
    """ux + vy = 0 exactly and vx - uy = vorticity of the uw kernel (to float rounding)."""
    sx, sy, ss, sr = W.onbody_random(400, seed=5)
    tx, ty, _, tr = W.onbody_random(64, seed=6)
    u, v, ux, vx, uy, vy = port.points_on_points_grad(core, sx, sy, ss, sr, tx, ty, tr)
    np.testing.assert_array_equal(ux, -vy)
    _, _, w = port.points_on_points(core, oracle.VELANDVORT, sx, sy, ss, sr, tx, ty, tr)
    assert W.rel_err(vx - uy, w) < 1e-5


def test_closed_vortex_ring(port):
    """Closed ring of equal-strength vortex panels: zero velocity inside, Gamma/r outside."""
    nx, ny, idx, vs, _ = W.circle_panels(400, radius=0.5)
    vs[:] = 1.0
    tx = np.array([0.1, -0.2, 0.0, 2.0], np.float32); ty = np.array([0.05, 0.1, 0.0, 0.0], np.float32)
    u, v = port.panels_on_points(nx, ny, idx, vs, None, tx, ty)
    assert np.abs(u[:3]).max() < 1e-5 and np.abs(v[:3]).max() < 1e-5
    gamma = 400 * 2 * 0.5 * math.sin(math.pi / 400)   # polygon perimeter x unit strength
    assert abs(abs(v[3]) - gamma / 2.0) < 1e-4 and abs(u[3]) < 1e-5
