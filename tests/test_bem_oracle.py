"""CPU: BEM coefficient / RHS restatement (oracle/o2d_oracle.c) against the reference-generated golden
matrices (tests/golden/ref_golden_bem.npz).  Tolerance: max |d| <= tol * max |entry of the assembled matrix|;
every block is part of one BEM matrix whose largest entries are the self-influence diagonal, 1/2, so off-diagonal
blocks between distant surfaces (entries ~1e-3, where the float32 reference formulation itself is only good to
~5e-9 absolute) are measured on that scale too."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gb():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_bem.npz"))


def surf(g, s):
    return (g[s + "x"], g[s + "y"], g[s + "idx"])


def mat_err(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 0.5))


@pytest.mark.parametrize("pair", ["aa", "bb", "cc", "ab", "ac", "ba", "bc", "ca", "cb"])
def test_coeff_port_vs_reference_golden(pair, gb, port):
    s, t = pair
    got = port.panels_on_panels_coeff(surf(gb, s), None if s == t else surf(gb, t))
    want = gb[f"coeff_{pair}"]
    assert got.shape == want.shape
    assert mat_err(got, want) < 2e-6
    if s == t:   # self influence: 1/2 on the vortex-tangential and source-normal diagonal, 0 on the cross terms
        n = got.shape[0] // 2
        d = np.arange(n)
        np.testing.assert_allclose(got[2 * d, 2 * d], 0.5, rtol=1e-7)
        np.testing.assert_allclose(got[2 * d + 1, 2 * d + 1], 0.5, rtol=1e-7)
        assert not got[2 * d, 2 * d + 1].any() and not got[2 * d + 1, 2 * d].any()


@pytest.mark.parametrize("pair", ["aa", "ac", "cb"])
def test_coeff_vortex_only_golden(pair, gb, port):
    s, t = pair
    got = port.panels_on_panels_coeff(surf(gb, s), None if s == t else surf(gb, t), both=False)
    assert mat_err(got, gb[f"coeffv_{pair}"]) < 2e-6


def test_bases_and_rhs_golden(gb, port):
    for s in "abc":
        got = np.stack(port.panel_bases(*surf(gb, s)))
        np.testing.assert_allclose(got, gb[f"bases_{s}"], atol=2e-7)
    rhs = port.vels_to_rhs_panels(gb["ax"], gb["ay"], gb["aidx"], gb["pu"], gb["pv"], gb["nbc"])
    np.testing.assert_allclose(rhs, gb["rhs_a"], atol=1e-6)


def test_leading_dimension_leaves_padding_untouched(gb, port):
    got = port.panels_on_panels_coeff(surf(gb, "b"), None, ld=2 * 61 + 1)
    assert got.shape == (122, 123) and not got[:, 122].any()
    np.testing.assert_array_equal(got[:, :122], port.panels_on_panels_coeff(surf(gb, "b")))


def test_uniform_flow_is_solved_by_the_matrix(gb, port):
    """Physics check of the assembled system: for a closed body in a uniform stream the vortex-sheet solution of
    the tangential-condition rows cancels the interior flow -- solve the vortex-only system with numpy and
    check that the circulation-free solution reproduces gamma = -2 U sin(theta) on the circle (potential flow)."""
    ax, ay, aidx = surf(gb, "a")
    A = port.panels_on_panels_coeff((ax, ay, aidx), None, both=False).T.astype(np.float64)   # rows = targets
    ttx, tty, _, _, area = port.panel_bases(ax, ay, aidx)
    rhs = -(1.0 * ttx + 0.0 * tty).astype(np.float64)            # vels_to_rhs_panels with u = freestream (1, 0)
    # augment with the zero-total-circulation row, least squares (what the reference's augmentation does)
    Aug = np.vstack([A, area[None, :].astype(np.float64)])
    gam = np.linalg.lstsq(Aug, np.concatenate([rhs, [0.0]]), rcond=None)[0]
    i0, i1 = aidx[0::2], aidx[1::2]
    th = np.arctan2(0.5 * (ay[i0] + ay[i1]), 0.5 * (ax[i0] + ax[i1]))
    # surface speed of potential flow past a circle is 2 U |sin(theta)|; the sheet strength equals it in magnitude
    assert np.abs(np.abs(gam) - np.abs(2.0 * np.sin(th))).max() < 2e-2


def test_rot_strengths_golden(gb, port):
    """Surfaces::add_rot_strengths_base: the reference's own body-bound Surfaces (geometric centre included)."""
    cx, cy, cidx = surf(gb, "c")
    area = gb["bases_c"][4]
    vs, ss = port.add_rot_strengths(cx, cy, cidx, area, tuple(gb["rot_center"]), float(gb["rot_rate"]), gb["rot_vs0"], gb["rot_ss0"])
    np.testing.assert_allclose(vs, gb["rot_vs"], atol=3e-7)
    np.testing.assert_allclose(ss, gb["rot_ss"], atol=3e-7)
    # adding with the opposite factor restores the strengths (find_vels: +1 before, -1 after the sums)
    vs2, ss2 = port.add_rot_strengths(cx, cy, cidx, area, tuple(gb["rot_center"]), -float(gb["rot_rate"]), vs, ss)
    np.testing.assert_allclose(vs2, gb["rot_vs0"], atol=3e-7)
    np.testing.assert_allclose(ss2, gb["rot_ss0"], atol=3e-7)
