"""CPU: the C restatement against the reference-generated BASELINE-size fixtures (tests/golden/ref_golden_large.npz),
and the fixture's own statement of the panel tolerance question."""
import os

import numpy as np
import pytest

import oracle
from omega2d_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def large():
    return np.load(os.path.join(HERE, "golden", "ref_golden_large.npz"))


def test_port_matches_reference_at_2m(port, large):
    oracle.set_threads()
    arrs = W.gaussian_blob(2_097_152, seed=2024)
    got = np.array([a.astype(np.float64).sum() for a in arrs])
    np.testing.assert_allclose(got, large["m2_input_checksum"], rtol=1e-12)
    tx, ty, tr = large["m2_tx"][:64], large["m2_ty"][:64], large["m2_tr"][:64]
    for core in ("wl", "v2"):
        u, v = port.points_on_points(core, oracle.VELONLY, *arrs, tx, ty, tr)
        ref = large[f"m2_{core}"][:, :64]
        assert W.rel_err((u, v), tuple(ref)) < 1e-6


def test_reference_panel_term_is_noisier_than_the_tolerance(large):
    """The documented exception, pinned: for the C4 panel set the reference's float32 panel evaluation is itself further
    than 1e-5 from float64 (random strengths: an ill-conditioned sum, 3e-3; smooth BEM-like strengths: 3.5e-5)."""
    for tag, lo in (("random", 1e-4), ("smooth", 1e-5)):
        noise = W.rel_err(tuple(large[f"c4_{tag}_panels_ref"]), tuple(large[f"c4_{tag}_panels_f64"]))
        assert noise > lo, (tag, noise)
    # ... while the port (same float32 formulas, same libm) reproduces the reference's numbers
    nx, ny, idx, _, _ = W.multi_circle_panels(4, 5000)
    p = oracle.port()
    tx, ty = large["c4_tx"][:128], large["c4_ty"][:128]
    got = p.panels_on_points(nx, ny, idx, large["c4_smooth_vs"], large["c4_smooth_ss"], tx, ty)
    ref = large["c4_smooth_panels_ref"][:, :128]
    exact = large["c4_smooth_panels_f64"][:, :128]
    assert W.rel_err(got, tuple(ref)) <= max(1e-5, 1.5 * W.rel_err(tuple(ref), tuple(exact)))
