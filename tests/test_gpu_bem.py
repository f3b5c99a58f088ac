"""GPU: o2d_panels_on_panels_coeff (CUDA, through the C ABI) against the reference-generated golden matrices and
the oracle.  Tolerance 1e-5 of the assembled matrix's largest entries (the 1/2 diagonal), see test_bem_oracle.py."""
import os

import numpy as np
import pytest

import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
from test_bem_oracle import mat_err, surf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5


@pytest.fixture(scope="module")
def gb():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_bem.npz"))


@pytest.mark.parametrize("pair", ["aa", "bb", "cc", "ab", "ac", "ba", "bc", "ca", "cb"])
def test_coeff_golden(pair, gb, ctx):
    s, t = pair
    got = ctx.panels_on_panels_coeff_np(surf(gb, s), gb[f"bases_{s}"][4], surf(gb, t), tuple(gb[f"bases_{t}"]),
                                        both=True, same=(s == t))
    want = gb[f"coeff_{pair}"]
    assert got.shape == want.shape
    assert mat_err(got, want) < TOL
    if s == t:
        d = np.arange(got.shape[0] // 2)
        assert (got[2 * d, 2 * d] == np.float32(0.5)).all() and (got[2 * d + 1, 2 * d + 1] == np.float32(0.5)).all()
        assert not got[2 * d, 2 * d + 1].any() and not got[2 * d + 1, 2 * d].any()


@pytest.mark.parametrize("pair", ["aa", "ac", "cb"])
def test_coeff_vortex_only_golden(pair, gb, ctx):
    s, t = pair
    got = ctx.panels_on_panels_coeff_np(surf(gb, s), gb[f"bases_{s}"][4], surf(gb, t), tuple(gb[f"bases_{t}"]),
                                        both=False, same=(s == t))
    assert mat_err(got, gb[f"coeffv_{pair}"]) < TOL


def test_coeff_augmented_leading_dimension(gb, ctx):
    """ld = rows + 1: the block lands inside an augmented matrix, the extra row is left for the caller."""
    b = surf(gb, "b")
    bases = tuple(gb["bases_b"])
    plain = ctx.panels_on_panels_coeff_np(b, bases[4], b, bases, both=True, same=True)
    aug = ctx.panels_on_panels_coeff_np(b, bases[4], b, bases, both=True, same=True, ld=2 * 61 + 1)
    assert aug.shape == (122, 123) and not aug[:, 122].any()
    np.testing.assert_array_equal(aug[:, :122], plain)


@pytest.mark.parametrize("npan", [1, 2, 129, 5000])
def test_coeff_sizes_vs_oracle(npan, ctx, port):
    nx, ny, idx, _, _ = W.circle_panels(max(npan, 3), radius=0.5)
    idx = idx[:2 * npan]
    s = (nx, ny, idx)
    bases = port.panel_bases(nx, ny, idx)
    want = port.panels_on_panels_coeff(s)
    got = ctx.panels_on_panels_coeff_np(s, bases[4], s, bases, both=True, same=True)
    assert mat_err(got, want) < TOL


def test_coeff_device_memory_path(gb, ctx):
    import torch
    a = surf(gb, "a"); bases = gb["bases_a"]
    host = ctx.panels_on_panels_coeff_np(a, bases[4], a, tuple(bases), both=True, same=True)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (a[0], a[1], bases[4], bases[0], bases[1], bases[2], bases[3])]
    didx = torch.from_numpy(a[2].astype(np.int32)).to(dev)
    c = torch.zeros((286, 286), device=dev)
    ctx.panels_on_panels_coeff(o2d.MEM_DEVICE, 143, 143, d[0], d[1], didx, d[2], 143, 143, d[0], d[1], didx, d[2],
                               d[3], d[4], d[5], d[6], 1, 1, 286, c)
    ctx.synchronize()
    np.testing.assert_array_equal(c.cpu().numpy(), host)
    with pytest.raises(o2d.O2DError):   # odd leading dimension is rejected for device memory
        ctx.panels_on_panels_coeff(o2d.MEM_DEVICE, 143, 143, d[0], d[1], didx, d[2], 143, 143, d[0], d[1], didx, d[2],
                                   d[3], d[4], d[5], d[6], 1, 1, 287, c)


def test_rot_strengths(gb, ctx):
    import torch
    cx, cy, cidx = surf(gb, "c")
    area = np.ascontiguousarray(gb["bases_c"][4])
    ctr, rate = tuple(gb["rot_center"]), float(gb["rot_rate"])
    vs, ss = gb["rot_vs0"].copy(), gb["rot_ss0"].copy()
    ctx.add_rot_strengths(o2d.MEM_HOST, 97, 97, cx, cy, cidx, area, ctr, rate, vs, ss)
    np.testing.assert_allclose(vs, gb["rot_vs"], atol=3e-7)
    np.testing.assert_allclose(ss, gb["rot_ss"], atol=3e-7)
    dev = torch.device("cuda:0")
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (cx, cy, area, gb["rot_vs0"], gb["rot_ss0"])]
    didx = torch.from_numpy(cidx.astype(np.int32)).to(dev)
    ctx.add_rot_strengths(o2d.MEM_DEVICE, 97, 97, d[0], d[1], didx, d[2], ctr, rate, d[3], d[4])
    ctx.synchronize()
    np.testing.assert_allclose(d[3].cpu().numpy(), gb["rot_vs"], atol=3e-7)
    np.testing.assert_allclose(d[4].cpu().numpy(), gb["rot_ss"], atol=3e-7)
