"""GPU: the one-process device group (o2d_group_*).  On a one-GPU box the group is built from several contexts on
device 0, which exercises the same block cutting, per-device host threads and slice write-back as several GPUs
(tools/group_check.py runs it on real multi-GPU boxes).  Results must equal the single-context call bit for bit: every
target is summed by the same kernel over the same sources."""
import numpy as np
import pytest

import omega2d_b200 as o2d
from omega2d_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def grp():
    g = o2d.Group(devices=[0, 0, 0], min_pairs=1_000_000)
    yield g
    g.close()


@pytest.mark.parametrize("core,results", [("rm", o2d.VELONLY), ("v2", o2d.VELANDVORT), ("wl", o2d.VELANDGRAD), ("exp", o2d.VELANDSHEAR)])
def test_points_on_points_equals_single_device(core, results, grp, ctx):
    assert grp.size == 3
    sx, sy, ss, sr = W.gaussian_blob(6001, seed=3)
    tx, ty, _, tr = W.gaussian_blob(5003, seed=4)          # 3 blocks of 2048, 2048, 907
    nout = {o2d.VELONLY: 2, o2d.VELANDVORT: 3, o2d.VELANDSHEAR: 3, o2d.VELANDGRAD: 6}[results]
    init = [np.full(tx.size, 0.25 * (k + 1), np.float32) for k in range(nout)]
    a = [x.copy() for x in init]; b = [x.copy() for x in init]
    if results == o2d.VELANDGRAD:
        grp.points_on_points(core, results, sx, sy, ss, sr, tx, ty, tr, a[0], a[1], None, *a[2:])
        ctx.points_on_points(core, results, o2d.MEM_HOST, sx.size, sx, sy, ss, sr, tx.size, tx, ty, tr, b[0], b[1], None, *b[2:])
    else:
        grp.points_on_points(core, results, sx, sy, ss, sr, tx, ty, tr, a[0], a[1], a[2] if nout == 3 else None)
        ctx.points_on_points(core, results, o2d.MEM_HOST, sx.size, sx, sy, ss, sr, tx.size, tx, ty, tr, b[0], b[1],
                             b[2] if nout == 3 else None, None, None, None, None)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


def test_find_vels_and_panels_equal_single_device(grp, ctx, port):
    vort = W.annulus_blobs(5000)
    panels = W.circle_panels(143)
    tx, ty, _, tr = W.annulus_blobs(4100, seed=8)
    u, v = np.empty(tx.size, np.float32), np.empty(tx.size, np.float32)
    grp.find_vels("v2", o2d.VELONLY, vort, panels, tx, ty, tr, u, v, fs=(1.0, 0.0))
    want = ctx.find_vels_np("v2", o2d.VELONLY, vort, panels, tx, ty, tr, (1.0, 0.0))
    assert np.array_equal(u, want[0]) and np.array_equal(v, want[1])
    assert W.rel_err((u, v), port.find_vels("v2", o2d.VELONLY, vort, panels, tx, ty, tr, (1.0, 0.0))[:2]) < 1e-5
    # panels on points alone, accumulate semantics
    a = [np.ones(tx.size, np.float32), np.ones(tx.size, np.float32)]
    grp.panels_on_points(panels, tx, ty, a[0], a[1])
    b = ctx.panels_on_points_np(*panels, tx, ty, init=(np.ones(tx.size, np.float32), np.ones(tx.size, np.float32)))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_points_on_panels_sources_cut(grp, ctx, port):
    sx, sy, ss, _ = W.annulus_blobs(7000)
    nx, ny, idx, _, _ = W.circle_panels(143)
    area = port.panel_areas(nx, ny, idx)
    pu, pv = np.full(143, 0.5, np.float32), np.full(143, -0.5, np.float32)
    grp.points_on_panels(sx, sy, ss, nx, ny, idx, area, pu, pv)
    want = ctx.points_on_panels_np(sx, sy, ss, nx, ny, idx, area, init=(np.full(143, 0.5, np.float32), np.full(143, -0.5, np.float32)))
    # the per-block sums are rounded to float before the host adds them: a few ulp, not bit-identical
    assert W.rel_err((pu, pv), want) < 1e-6
    assert W.rel_err((pu, pv), port.points_on_panels(sx, sy, ss, nx, ny, idx, area, pu=np.full(143, 0.5, np.float32), pv=np.full(143, -0.5, np.float32))[:2]) < 1e-5


def test_small_problem_stays_on_one_device_and_errors_propagate(ctx):
    g = o2d.Group(devices=[0, 0])          # default threshold: 2e9 pairs per device
    sx, sy, ss, sr = W.gaussian_blob(300, seed=1)
    u, v = np.zeros(300, np.float32), np.zeros(300, np.float32)
    g.points_on_points("rm", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr, u, v)
    want = ctx.points_on_points_np("rm", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr)
    assert np.array_equal(u, want[0])
    with pytest.raises(o2d.O2DError) as ei:
        g.points_on_points("rm", o2d.PSIONLY, sx, sy, ss, sr, sx, sy, sr, u, v)
    assert ei.value.code == o2d.ERR_UNSUPPORTED
    g.close()
    g2 = o2d.Group(devices=[0, 0], min_pairs=1)
    with pytest.raises(o2d.O2DError) as ei:     # failure inside the worker threads reaches the caller with its message
        g2.points_on_points("rm", o2d.VELANDVORT, sx, sy, ss, sr, np.tile(sx, 10), np.tile(sy, 10), np.tile(sr, 10),
                            np.zeros(3000, np.float32), np.zeros(3000, np.float32), None)
    assert ei.value.code == o2d.ERR_INVALID and "tq" in str(ei.value)
    g2.close()
    with pytest.raises(o2d.O2DError):
        o2d.Group(devices=[99])
