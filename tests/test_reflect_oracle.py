"""CPU: clear_inner_layer (push method) and reflect_interior restatements against the reference-generated golden vectors, plus the
properties the operation must have whatever the rounding: idempotence, nothing left inside the cutoff layer,
untouched particles bit-identical."""
import math
import os

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POS_TOL = 1e-5     # of the body size (1.0): north_star's relative tolerance applied to positions


@pytest.fixture(scope="module")
def gr():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_reflect.npz"))


def check_cleared(fn, g, tag, want_key):
    nx, ny, idx = g[f"{tag}_nx"], g[f"{tag}_ny"], g[f"{tag}_idx"]
    px, py = g[f"{tag}_px"], g[f"{tag}_py"]
    cf, ips = float(g["cf"]), float(g["ips"])
    x, y, moved = fn(nx, ny, idx, px, py, cf, ips)
    want = g[want_key]
    d = np.hypot(x.astype(np.float64) - want[0], y.astype(np.float64) - want[1])
    assert d.max() < POS_TOL, d.max()
    untouched = (want[0] == px) & (want[1] == py)
    assert untouched.sum() > 0 and (x[untouched] == px[untouched]).all() and (y[untouched] == py[untouched]).all()
    assert moved >= (~untouched).sum()          # corners may take more than one push
    # idempotent: a second pass moves nothing
    x2, y2, moved2 = fn(nx, ny, idx, x, y, cf, ips)
    assert moved2 == 0 and (x2 == x).all() and (y2 == y).all()
    return x, y


def test_circle_golden(gr, port):
    x, y = check_cleared(port.clear_inner_layer, gr, "circ", "circ_out")
    # everything ends up outside the cutoff layer around the 143-gon (sagitta of a panel allowed for)
    cf, ips = float(gr["cf"]), float(gr["ips"])
    rmin = 0.5 * math.cos(math.pi / 143) + cf * ips - 0.0011 * ips
    assert np.hypot(x, y).min() > rmin - 1e-6
    check_cleared(port.clear_inner_layer, gr, "circ", "circ_out_fldpt")


def test_square_with_corners_golden(gr, port):
    x, y = check_cleared(port.clear_inner_layer, gr, "sq", "sq_out")
    cf, ips = float(gr["cf"]), float(gr["ips"])
    # Chebyshev distance from the square's boundary is at least the layer thickness (outside) for all points
    lay = cf * ips - 0.0011 * ips
    cheb = np.maximum(np.abs(x), np.abs(y))
    assert cheb.min() > 0.5 + lay / math.sqrt(2) - 1e-6


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_reference_fresh(port):
    from omega2d_b200 import workloads as W
    R = oracle.ref("v2")
    nx, ny, idx, _, _ = W.multi_circle_panels(3, 200, radius=0.3, spacing=0.9)
    rng = np.random.default_rng(9)
    x = rng.uniform(-1.4, 1.4, 20000).astype(np.float32); y = rng.uniform(1.1, 1.9, 20000).astype(np.float32)
    a = R.clear_inner_layer(nx, ny, idx, x, y, 0.2, 0.05)
    b = port.clear_inner_layer(nx, ny, idx, x, y, 0.2, 0.05)
    assert np.hypot(a[0] - b[0], a[1] - b[1]).max() < POS_TOL
    assert b[2] > 1000


def check_reflected(fn, g, tag, kx, ky, want_key, body_test):
    """reflect_interior: golden parity; untouched particles bit-identical; every particle that started strictly
    inside ends up outside; a second pass moves (almost) nothing."""
    nx, ny, idx = g[f"{tag}_nx"], g[f"{tag}_ny"], g[f"{tag}_idx"]
    px, py = g[kx], g[ky]
    x, y, moved = fn(nx, ny, idx, px, py)
    want = g[want_key]
    d = np.hypot(x.astype(np.float64) - want[0], y.astype(np.float64) - want[1])
    assert d.max() < POS_TOL, d.max()
    changed = (want[0] != px) | (want[1] != py)
    assert (x[~changed] == px[~changed]).all() and (y[~changed] == py[~changed]).all()
    assert moved >= changed.sum() and moved <= changed.sum() + 4    # a reflection can land on the same float
    inside = body_test(px, py)
    assert inside.sum() > 100 and (~body_test(x[inside], y[inside])).all()
    return x, y


def inside_circle(x, y):      # strictly inside the inscribed circle of the 143-gon
    return np.hypot(x, y) < 0.5 * math.cos(math.pi / 143) - 1e-4


def inside_square(x, y):
    return (np.maximum(np.abs(x), np.abs(y)) < 0.5 - 1e-4) & ((x != 0) | (y != 0))


def test_reflect_circle_golden(gr, port):
    check_reflected(port.reflect_interior, gr, "circ", "circ_px", "circ_py", "circ_reflect", inside_circle)


def test_reflect_square_ties_golden(gr, port):
    x, y = check_reflected(port.reflect_interior, gr, "sq", "sq_rx", "sq_ry", "sq_reflect", inside_square)
    # the centre: eight tied hits whose normals cancel -> the reference falls back to hits[0] and still divides
    # the contact point by the hit count (src/Reflect.h:305-327): (0,0) -> (-0.5/8 - 0.5, 0)
    assert (x[0], y[0]) == (-0.5625, 0.0)
    # mirror images: distance to the wall is preserved
    assert (x[1], y[1]) == (0.0, 0.75) and (x[11], y[11]) == (np.float32(0.51), 0.0)


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_reflect_port_matches_reference_fresh(port):
    from omega2d_b200 import workloads as W
    R = oracle.ref("v2")
    nx, ny, idx, _, _ = W.multi_circle_panels(3, 200, radius=0.3, spacing=0.9)
    rng = np.random.default_rng(19)
    x = rng.uniform(-1.4, 1.4, 20000).astype(np.float32); y = rng.uniform(1.1, 1.9, 20000).astype(np.float32)
    one = np.ones_like(x)
    a = R.reflect_interior(nx, ny, idx, x, y, one, 0.05 * one)
    b = port.reflect_interior(nx, ny, idx, x, y)
    assert np.hypot(a[0] - b[0], a[1] - b[1]).max() < POS_TOL
    assert b[2] > 1000
