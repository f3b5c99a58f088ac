"""CPU: the C restatement against the compiled UNMODIFIED reference (oracle/_ref) on fresh
random inputs.  Runs wherever oracle/_ref exists (built in the container that has
/root/reference; the prebuilt libraries travel to the GPU box)."""
import numpy as np
import pytest

import oracle
from omega2d_b200 import workloads as W

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("core", oracle.CORES)
def test_port_matches_reference(core, port):
    R = oracle.ref(core)
    sx, sy, ss, sr = W.onbody_random(1500, seed=31)
    tx, ty, _, tr = W.onbody_random(300, seed=32)
    for res in (oracle.VELONLY, oracle.VELANDVORT, oracle.VELANDSHEAR):
        for trr in (tr, None):
            a = R.points_on_points(res, sx, sy, ss, sr, tx, ty, trr)
            b = port.points_on_points(core, res, sx, sy, ss, sr, tx, ty, trr)
            assert W.rel_err(b[:2], a[:2]) < 1e-6
            if len(a) == 3:
                assert W.rel_err(b[2], a[2]) < 3e-6
    a = R.points_on_points_grad(sx, sy, ss, sr, tx, ty, tr)
    b = port.points_on_points_grad(core, sx, sy, ss, sr, tx, ty, tr)
    assert W.rel_err(b[2:], a[2:]) < 3e-6


def test_reference_self_interaction_is_included(port):
    """i==j pairs are part of the sum: 0 for u,v, core(0) for vorticity."""
    R = oracle.ref("v2")
    x = np.array([0.3], np.float32); y = np.array([-0.2], np.float32)
    s = np.array([2.0], np.float32); r = np.array([0.1], np.float32)
    u, v, w = R.points_on_points(oracle.VELANDVORT, x, y, s, r, x, y, r)
    assert u[0] == 0 and v[0] == 0 and w[0] > 0
    u2, v2, w2 = port.points_on_points("v2", oracle.VELANDVORT, x, y, s, r, x, y, r)
    assert w2[0] == w[0]


def test_panels_and_vorticity_match_reference(port):
    R = oracle.ref("v2")
    nx, ny, idx, vs, ss = W.circle_panels(64, seed=3)
    x, y, s, r = W.annulus_blobs(500, seed=4)
    assert W.rel_err(port.panels_on_points(nx, ny, idx, vs, ss, x, y), R.panels_on_points(nx, ny, idx, vs, ss, x, y)) < 3e-6
    assert W.rel_err(port.panels_on_points(nx, ny, idx, vs, None, x, y), R.panels_on_points(nx, ny, idx, vs, None, x, y)) < 3e-6
    pu, pv, area = R.points_on_panels(x, y, s, nx, ny, idx)
    assert W.rel_err(port.points_on_panels(x, y, s, nx, ny, idx, area), (pu, pv)) < 5e-6
    assert W.rel_err(port.points_vorticity(x, y, s, r, x[:100], y[:100]), R.points_vorticity(x, y, s, r, x[:100], y[:100])) < 1e-6
