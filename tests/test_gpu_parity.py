"""GPU: the CUDA path, called through the C ABI, against the oracle.

Tolerance is BASELINE.json's: max_i ||du_i|| / max_i ||u_i|| <= 1e-5 (float32), applied to
velocities, vorticity, shear and every gradient component.  Oracle = the C restatement
(pinned to the reference by the CPU tests) and the committed reference golden vectors."""
import numpy as np
import pytest

import omega2d_b200 as o2d
import oracle
from omega2d_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-5
RES = {"u": o2d.VELONLY, "uw": o2d.VELANDVORT, "ue": o2d.VELANDSHEAR}


@pytest.mark.parametrize("core", o2d.CORES)
@pytest.mark.parametrize("blob", [0, 1])
def test_golden_vectors(core, blob, ctx, golden):
    g = golden
    tr = g["tr"] if blob else None
    got = ctx.points_on_points_np(core, o2d.VELONLY, g["sx"], g["sy"], g["ss"], g["sr"], g["tx"], g["ty"], tr)
    assert W.rel_err(got, tuple(g[f"{core}_u_{blob}"])) < TOL
    sl = slice(40, None) if core == "exp" else slice(None)
    trs = None if tr is None else tr[sl]
    for key in ("uw", "ue"):
        got = ctx.points_on_points_np(core, RES[key], g["sx"], g["sy"], g["ss"], g["sr"], g["tx"][sl], g["ty"][sl], trs)
        want = g[f"{core}_{key}_{blob}"]
        assert W.rel_err(got[:2], (want[0], want[1])) < TOL
        assert W.rel_err(got[2], want[2]) < TOL, (core, key, blob)
    if blob:
        got = ctx.points_on_points_np(core, o2d.VELANDGRAD, g["sx"], g["sy"], g["ss"], g["sr"],
                                      g["tx"][sl], g["ty"][sl], g["tr"][sl])
        want = g[f"{core}_ug_1"]
        assert W.rel_err(got[:2], (want[0], want[1])) < TOL
        for k in range(2, 6):
            assert W.rel_err(got[k], want[k]) < TOL, (core, "grad", k)
        np.testing.assert_array_equal(got[2], -got[5])   # ux + vy == 0 exactly, as in the reference


@pytest.mark.parametrize("core", o2d.CORES)
@pytest.mark.parametrize("ns,nt", [(1, 1), (2, 3), (1023, 257), (1024, 1024), (1025, 1025), (5000, 2049), (20000, 777)])
def test_ragged_sizes_vs_oracle(core, ns, nt, ctx, port):
    sx, sy, ss, sr = W.onbody_random(ns, seed=ns)
    tx, ty, _, tr = W.onbody_random(nt, seed=nt + 1)
    want = port.points_on_points(core, oracle.VELONLY, sx, sy, ss, sr, tx, ty, tr)
    got = ctx.points_on_points_np(core, o2d.VELONLY, sx, sy, ss, sr, tx, ty, tr)
    assert W.rel_err(got, want) < TOL


@pytest.mark.parametrize("core", o2d.CORES)
@pytest.mark.parametrize("res", ["u", "uw", "ue", "ug"])
def test_uniform_and_mixed_target_radii(core, res, ctx, port):
    """The kernel switches, CTA by CTA, to a cheaper loop when all of a thread's targets share one
    radius (uniform-radius sets are the reference's normal case; inert targets likewise).  Cover
    (i) all-uniform targets incl. a ragged last CTA and (ii) a target set whose first part is
    uniform and whose tail is not, so both loops run inside one launch."""
    restype = {"u": oracle.VELONLY, "uw": oracle.VELANDVORT, "ue": oracle.VELANDSHEAR, "ug": oracle.VELANDGRAD}[res]
    mine = {"u": o2d.VELONLY, "uw": o2d.VELANDVORT, "ue": o2d.VELANDSHEAR, "ug": o2d.VELANDGRAD}[res]
    sx, sy, ss, sr = W.onbody_random(6000, seed=21)
    tx, ty, _, tr = W.onbody_random(3333, seed=22)
    tx = tx + np.float32(3e-4)   # no coincident pairs (EXP gradient terms are NaN there, tested elsewhere)
    for mode in ("uniform", "mixed"):
        trr = tr.copy()
        trr[: (3333 if mode == "uniform" else 2048)] = np.float32(0.0125)
        if res == "ug":
            want = port.points_on_points_grad(core, sx, sy, ss, sr, tx, ty, trr)
        else:
            want = port.points_on_points(core, restype, sx, sy, ss, sr, tx, ty, trr)
        got = ctx.points_on_points_np(core, mine, sx, sy, ss, sr, tx, ty, trr)
        assert W.rel_err(got[:2], want[:2]) < TOL, (mode, "vel")
        for k in range(2, len(got)):
            assert W.rel_err(got[k], want[k]) < TOL, (mode, k)


def test_empty_inputs_are_noops(ctx):
    e = np.zeros(0, np.float32)
    a = np.ones(5, np.float32)
    u, v = ctx.points_on_points_np("v2", o2d.VELONLY, e, e, e, e, a, a, a, init=(a * 3, a * 4))
    assert (u == 3).all() and (v == 4).all()
    u, v = ctx.points_on_points_np("v2", o2d.VELONLY, a, a, a, a, e, e, e)
    assert u.size == 0


def test_accumulate_semantics(ctx, golden):
    g = golden
    got = ctx.points_on_points_np("v2", o2d.VELONLY, g["sx"], g["sy"], g["ss"], g["sr"], g["tx"], g["ty"], g["tr"],
                                  init=(g["u0"], g["v0"]))
    assert W.rel_err(got, tuple(g["v2_u_1_accum"])) < TOL


def test_self_interaction_targets_are_sources(ctx, port):
    """C1 shape: targets == sources (self pairs included, contribute 0 to u,v and core(0) to w)."""
    x, y, s, r = W.c1_speedtest(4000)
    for core in ("v2", "rm", "wl"):
        want = port.points_on_points(core, oracle.VELANDVORT, x, y, s, r, x, y, r)
        got = ctx.points_on_points_np(core, o2d.VELANDVORT, x, y, s, r, x, y, r)
        assert W.rel_err(got[:2], want[:2]) < TOL
        assert W.rel_err(got[2], want[2]) < TOL


def test_exp_gradient_nan_matches_reference(ctx, golden):
    g = golden
    u, v, w = ctx.points_on_points_np("exp", o2d.VELANDVORT, g["sx"], g["sy"], g["ss"], g["sr"],
                                      g["tx"][:40], g["ty"][:40], g["tr"][:40])
    assert np.isfinite(u).all() and np.isfinite(v).all() and np.isnan(w).all()


def test_unsupported_and_invalid(ctx):
    a = np.ones(4, np.float32)
    with pytest.raises(o2d.O2DError) as ei:
        ctx.points_on_points("v2", o2d.PSIONLY, o2d.MEM_HOST, 4, a, a, a, a, 4, a, a, a, a.copy(), a.copy())
    assert ei.value.code == o2d.ERR_UNSUPPORTED
    with pytest.raises(o2d.O2DError) as ei:
        ctx.points_on_points(7, o2d.VELONLY, o2d.MEM_HOST, 4, a, a, a, a, 4, a, a, a, a.copy(), a.copy())
    assert ei.value.code == o2d.ERR_INVALID
    with pytest.raises(o2d.O2DError) as ei:   # vorticity requested without an output array
        ctx.points_on_points("v2", o2d.VELANDVORT, o2d.MEM_HOST, 4, a, a, a, a, 4, a, a, a, a.copy(), a.copy())
    assert ei.value.code == o2d.ERR_INVALID


def test_panels_golden(ctx, golden):
    g = golden
    got = ctx.panels_on_points_np(g["nx"], g["ny"], g["idx"], g["vs"], g["sps"], g["px"], g["py"])
    assert W.rel_err(got, tuple(g["pan_vs"])) < TOL
    got = ctx.panels_on_points_np(g["nx"], g["ny"], g["idx"], g["vs"], None, g["px"], g["py"])
    assert W.rel_err(got, tuple(g["pan_v"])) < TOL
    got = ctx.points_on_panels_np(g["px"], g["py"], g["ps"], g["nx"], g["ny"], g["idx"], g["area"])
    assert W.rel_err(got, tuple(g["ptpan"])) < TOL


def _panel_check(got, want, exact):
    """Within 1e-5 of the oracle -- unless the oracle's float32 formulation (difference of two
    atan2f, log of a ratio, reference src/Kernels.h:262-268,338) is itself further than that from
    exact (float64) arithmetic, which happens for short panels; then: within 1e-5 of exact and no
    further from the oracle than 1.5x the oracle's own error."""
    noise = W.rel_err(want, exact)
    assert W.rel_err(got, exact) < TOL, ("vs exact", W.rel_err(got, exact))
    assert W.rel_err(got, want) < max(TOL, 1.5 * noise), ("vs oracle", W.rel_err(got, want), "oracle noise", noise)


@pytest.mark.parametrize("npan,nt", [(1, 5), (143, 10000), (257, 300), (5000, 4096)])
def test_panels_vs_oracle(npan, nt, ctx, port):
    """Smooth sheet strengths (flow-over-cylinder like: vs ~ sin(theta)), the physically relevant case."""
    from oracle.f64 import panels_on_points_f64, points_on_panels_f64
    nx, ny, idx, vs, ss = W.circle_panels(max(npan, 3), radius=0.5, seed=npan)
    th = np.arctan2(ny, nx)
    vs = (2.0 * np.sin(th) + 0.3).astype(np.float32); ss = (0.5 * np.cos(th)).astype(np.float32)
    idx, vs, ss = idx[:2 * npan], vs[:npan], ss[:npan]
    x, y, s, r = W.annulus_blobs(nt, seed=nt)
    want = port.panels_on_points(nx, ny, idx, vs, ss, x, y)
    got = ctx.panels_on_points_np(nx, ny, idx, vs, ss, x, y)
    assert W.rel_err(got, want) < TOL
    area = port.panel_areas(nx, ny, idx)
    want = port.points_on_panels(x, y, s, nx, ny, idx, area)
    got = ctx.points_on_panels_np(x, y, s, nx, ny, idx, area)
    _panel_check(got, want, points_on_panels_f64(x, y, s, nx, ny, idx, area))


def test_panels_random_strengths_noise_floor(ctx, port):
    """5000 short panels with RANDOM +-1 strengths: the far field nearly cancels and the float32
    reference formulation carries rounding noise above 1e-5 of max|u| by itself (3.8e-5 here)."""
    from oracle.f64 import panels_on_points_f64
    nx, ny, idx, vs, ss = W.circle_panels(5000, radius=0.5, seed=5000)
    x, y, _, _ = W.annulus_blobs(2048, seed=4096)
    want = port.panels_on_points(nx, ny, idx, vs, ss, x, y)
    got = ctx.panels_on_points_np(nx, ny, idx, vs, ss, x, y)
    _panel_check(got, want, panels_on_points_f64(nx, ny, idx, vs, ss, x, y))


@pytest.mark.parametrize("angle", [0.0, 0.7, 2.3])
def test_panel_series_tiers_per_target(angle, ctx):
    """The far field is a Laurent series cut after 2, 3 or 4 terms by distance (23 / 6.2 / 3.2 panel lengths, per warp
    vote), the near field the general form.  Targets on rays from 0.05 to 3000 panel lengths from ONE short panel far
    from the origin, shuffled so warps mix tiers: every target within 1e-6 of float64 RELATIVE TO ITS OWN value (the
    reference's own float32 form is off by up to 4e-4 here, src/Kernels.h:312-405)."""
    from oracle.f64 import panels_on_points_f64, points_on_panels_f64
    L = 1.0e-3
    c = np.array([1.25, -0.75])
    d = 0.5 * L * np.array([np.cos(angle), np.sin(angle)])
    nx = np.array([c[0] - d[0], c[0] + d[0]], np.float32); ny = np.array([c[1] - d[1], c[1] + d[1]], np.float32)
    idx = np.array([0, 1], np.uint32)
    vs = np.array([0.8], np.float32); ss = np.array([-0.6], np.float32)
    rng = np.random.default_rng(5)
    dist = L * np.logspace(np.log10(0.05), np.log10(3000.0), 6000)
    th = rng.uniform(0, 2 * np.pi, dist.size)
    x = (c[0] + dist * np.cos(th)).astype(np.float32); y = (c[1] + dist * np.sin(th)).astype(np.float32)
    for order in (np.arange(dist.size), rng.permutation(dist.size)):     # sorted: uniform warps; shuffled: mixed warps
        xs, ys = x[order], y[order]
        got = ctx.panels_on_points_np(nx, ny, idx, vs, ss, xs, ys)
        exact = panels_on_points_f64(nx, ny, idx, vs, ss, xs, ys)
        mag = np.hypot(exact[0], exact[1])
        err = np.hypot(got[0] - exact[0], got[1] - exact[1]) / mag
        assert err.max() < 1e-6, (err.max(), dist[order][err.argmax()] / L)
    # targets exactly ON the panel between its nodes (ustar = +-pi): the side the reference's two-atan2 form picks
    if angle == 0.0:
        for sgn in (1.0, -1.0):
            nxs = (nx if sgn > 0 else nx[::-1]).copy()
            xs = np.array([c[0] + 0.2 * L, c[0] - 0.3 * L], np.float32); ys = np.array([ny[0], ny[0]], np.float32)
            got = ctx.panels_on_points_np(nxs, ny, idx, vs, ss, xs, ys)
            exact = panels_on_points_f64(nxs, ny, idx, vs, ss, xs, ys)
            assert W.rel_err(got, exact) < 1e-6, (sgn, got, exact)
    # the reciprocal kernel: unit-strength particles, one at a time, on the panel
    area = np.array([L], np.float32)
    for k in range(0, dist.size, 97):
        one = np.array([1.0], np.float32)
        got = ctx.points_on_panels_np(x[k:k + 1], y[k:k + 1], one, nx, ny, idx, area)
        exact = points_on_panels_f64(x[k:k + 1], y[k:k + 1], one, nx, ny, idx, area)
        mag = float(np.hypot(exact[0][0], exact[1][0]))
        assert float(np.hypot(got[0][0] - exact[0][0], got[1][0] - exact[1][0])) / mag < 1e-6, (k, dist[k] / L)


def test_closed_ring_property(ctx):
    nx, ny, idx, vs, _ = W.circle_panels(2000, radius=0.5)
    vs[:] = 1.0
    rng = np.random.default_rng(0)
    rho = 0.4 * np.sqrt(rng.uniform(0, 1, 5000)); th = rng.uniform(0, 6.2832, 5000)
    tx = (rho * np.cos(th)).astype(np.float32); ty = (rho * np.sin(th)).astype(np.float32)
    u, v = ctx.panels_on_points_np(nx, ny, idx, vs, None, tx, ty)
    assert np.abs(u).max() < 2e-5 and np.abs(v).max() < 2e-5


def test_vorticity(ctx, golden, port):
    g = golden
    got = ctx.points_vorticity_np(g["sx"], g["sy"], g["ss"], g["sr"], g["tx"], g["ty"])
    assert W.rel_err(got, g["vort"]) < TOL
    sx, sy, ss, sr = W.gaussian_blob(6000, seed=3)
    want = port.points_vorticity(sx, sy, ss, sr, sx[:1500], sy[:1500])
    got = ctx.points_vorticity_np(sx, sy, ss, sr, sx[:1500], sy[:1500])
    assert W.rel_err(got, want) < TOL


def test_finalize_and_zero(ctx, golden):
    g = golden
    uw = g["v2_uw_1"]
    u, v, w = uw[0].copy(), uw[1].copy(), uw[2].copy()
    ctx.finalize_vels(o2d.MEM_HOST, u.size, u, v, w, None, (1.0, -0.5))
    np.testing.assert_array_equal(np.stack([u, v, w]), g["fin"])
    ctx.zero_vels(o2d.MEM_HOST, u.size, u, v, w, None)
    assert not u.any() and not v.any() and not w.any()


def test_legacy_plugin_symbols(ctx, port):
    """external_vel_solver_f_/_tr_f_ as Omega2D calls them (reference src/Influence.h:82-88)."""
    sx, sy, ss, sr = W.onbody_random(3000, seed=8)
    tx, ty, _, tr = W.onbody_random(500, seed=9)
    u0 = np.full(500, 0.25, np.float32)
    tu, tv, flops = o2d.external_vel_solver(sx, sy, ss, sr, tx, ty, tr, tu=u0, tv=u0)
    want = port.points_on_points("v2", oracle.VELONLY, sx, sy, ss, sr, tx, ty, tr, u0, u0)
    assert W.rel_err((tu, tv), want) < TOL
    assert flops == pytest.approx(500 * (2.0 + 19.0 * 3000), rel=1e-6)
    tu, tv, flops = o2d.external_vel_solver(sx, sy, ss, sr, tx, ty, None)
    want = port.points_on_points("v2", oracle.VELONLY, sx, sy, ss, sr, tx, ty, None)
    assert W.rel_err((tu, tv), want) < TOL
    assert flops == pytest.approx(500 * (2.0 + 16.0 * 3000), rel=1e-6)


def test_device_memory_path_and_determinism(ctx):
    import torch
    sx, sy, ss, sr = W.gaussian_blob(50000, seed=5)
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(a).to(dev) for a in (sx, sy, ss, sr)]
    outs = []
    for _ in range(2):
        tu = torch.zeros(50000, device=dev); tv = torch.zeros(50000, device=dev)
        ctx.points_on_points("rm", o2d.VELONLY, o2d.MEM_DEVICE, 50000, *t, 50000, t[0], t[1], t[3], tu, tv)
        ctx.synchronize()
        outs.append((tu.cpu().numpy(), tv.cpu().numpy()))
    np.testing.assert_array_equal(outs[0][0], outs[1][0])   # bit-reproducible
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    host = ctx.points_on_points_np("rm", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr)
    np.testing.assert_array_equal(host[0], outs[0][0])      # host path == device path


def test_sharded_single_rank_equals_plain(ctx):
    """o2d_points_on_points_sharded with one rank (no communicator) takes the chunked
    local-block path and must equal the plain call (same pair terms; the sums are grouped by 1024-source tiles in
    the sharded path and by 256-source tiles in the plain path at this size, so not bit for bit)."""
    sx, sy, ss, sr = W.gaussian_blob(30000, seed=6)
    a = ctx.points_on_points_np("v2", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr)
    tu = np.zeros(30000, np.float32); tv = np.zeros(30000, np.float32)
    ctx.points_on_points_sharded("v2", o2d.VELONLY, o2d.MEM_HOST, 30000, 30000, sx, sy, ss, sr,
                                 30000, sx, sy, sr, tu, tv)
    assert W.rel_err((tu, tv), a) < 1e-6


# ---- device-resident find_vels / convection step -----------------------------------
@pytest.fixture(scope="module")
def gs():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return np.load(os.path.join(root, "tests", "golden", "ref_golden_substep.npz"))


@pytest.mark.parametrize("core", o2d.CORES)
def test_find_vels_golden(core, ctx, gs):
    g = gs
    fs = tuple(g["fs"])
    vort = (g["x"], g["y"], g["s"], g["r2"])
    pan = (g["nx"], g["ny"], g["idx"], g["pvs"], g["pss"])
    for res, key in ((o2d.VELONLY, "u"), (o2d.VELANDVORT, "uw"), (o2d.VELANDSHEAR, "ue")):
        for tr, tag in ((g["tr"], "blob"), (None, "inert")):
            got = ctx.find_vels_np(core, res, vort, pan, g["tx"], g["ty"], tr, fs)
            want = g[f"{core}_fv_{key}_{tag}"]
            assert W.rel_err(got[:2], (want[0], want[1])) < TOL
            if len(got) == 3:
                assert W.rel_err(got[2], want[2]) < TOL
    got = ctx.find_vels_np(core, o2d.VELONLY, vort, None, g["tx"], g["ty"], g["tr"], fs)
    assert W.rel_err(got, tuple(g[f"{core}_fv_u_nopanels"])) < TOL


@pytest.mark.parametrize("core", o2d.CORES)
def test_advect_golden(core, ctx, gs):
    from test_substep_oracle import check_step
    g = gs
    fs, dt = tuple(g["fs"]), float(g["dt"])
    for order in (1, 2, 3):
        for rk, tag in (("r", ""), ("r2", "_r2")):
            got = ctx.advect_np(core, order, g["x"], g["y"], g["s"], g[rk], dt, None, fs)
            check_step(got, g[f"{core}_adv{order}{tag}"], g["x"], g["y"], TOL)
    pan = (g["nx"], g["ny"], g["idx"], g["pvs"], g["pss"])
    got = ctx.advect_np(core, 2, g["x"], g["y"], g["s"], g["r2"], dt, pan, fs)
    check_step(got, g[f"{core}_adv2_panels"], g["x"], g["y"], TOL)


def test_advect_c1_speedtest_vs_oracle(ctx, port):
    """C1 (speedtest20kx10 restated): one RK2 step of 20k uniform-radius blobs, V2 core."""
    from test_substep_oracle import check_step
    x, y, s, r = W.c1_speedtest(20000)
    want = port.advect("v2", 2, x, y, s, r, 0.005)
    got = ctx.advect_np("v2", 2, x, y, s, r, 0.005)
    check_step(got, want, x, y, TOL)


def test_advect_steps_equals_repeated_steps_and_tracks_the_oracle(ctx, port):
    """o2d_advect_steps keeps the particles on the device between steps: same bits as step-by-step calls; ten RK2 steps
    of the C1 speedtest (what examples/speedtest20kx10.json runs) stay within tolerance of the CPU restatement."""
    x, y, s, r = W.c1_speedtest(20000)
    gx, gy = x.copy(), y.copy()
    u, v = np.zeros_like(x), np.zeros_like(x)
    ctx.advect("v2", 2, o2d.MEM_HOST, x.size, gx, gy, s, r, u, v, dt=0.005, nsteps=10)
    sx_, sy_ = x.copy(), y.copy()
    su, sv = np.zeros_like(x), np.zeros_like(x)
    wx, wy = x, y
    for k in range(10):
        # every single step, started from the ORACLE's state of that step, is held to the plain tolerance: both the
        # velocities the step leaves behind and the displacement it makes (this is the 1e-5 check of the run; the
        # accumulated positions below drift apart only because close pairs amplify last-bit differences chaotically)
        px, py = wx.copy(), wy.copy()
        pu, pv = np.zeros_like(x), np.zeros_like(x)
        ctx.advect("v2", 2, o2d.MEM_HOST, x.size, px, py, s, r, pu, pv, dt=0.005)
        ox, oy = wx, wy
        ctx.advect("v2", 2, o2d.MEM_HOST, x.size, sx_, sy_, s, r, su, sv, dt=0.005)
        wx, wy, wu, wv = port.advect("v2", 2, wx, wy, s, r, 0.005)
        assert W.rel_err((pu, pv), (wu, wv)) < TOL, k
        assert W.rel_err((px - ox, py - oy), (wx - ox, wy - oy)) < TOL, k
    assert np.array_equal(gx, sx_) and np.array_equal(gy, sy_) and np.array_equal(u, su) and np.array_equal(v, sv)
    # 20 000 strong vortices (circulations ~ U[-1,1]) at random: close pairs orbit each other fast and amplify the
    # last-bit differences of every evaluation, so after ten steps the bulk must agree tightly, the worst pair loosely
    d = np.hypot(gx - wx, gy - wy)
    assert np.percentile(d, 90) < TOL * 4.0 and d.max() < 5e-3, (np.percentile(d, 90), d.max())   # domain is [-2,2]^2
    assert np.hypot(gx - x, gy - y).max() > 1e-3                  # and the particles did move


def test_advect_device_memory_equals_host_path(ctx):
    import torch
    x, y, s, r = W.c1_speedtest(5000, seed=3)
    host = ctx.advect_np("rm", 3, x, y, s, r, 0.01, None, (0.3, 0.0))
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(a.copy()).to(dev) for a in (x, y, s, r)]
    u = torch.empty(5000, device=dev); v = torch.empty(5000, device=dev)
    ctx.advect("rm", 3, o2d.MEM_DEVICE, 5000, t[0], t[1], t[2], t[3], u, v, fs=(0.3, 0.0), dt=0.01)
    ctx.synchronize()
    for a, b in zip(host, (t[0], t[1], u, v)):
        np.testing.assert_array_equal(a, b.cpu().numpy())


def test_find_vels_and_advect_argument_errors(ctx):
    a = np.ones(4, np.float32)
    with pytest.raises(o2d.O2DError) as ei:
        ctx.advect("v2", 4, o2d.MEM_HOST, 4, a.copy(), a.copy(), a, a, a.copy(), a.copy(), dt=0.1)
    assert ei.value.code == o2d.ERR_INVALID
    with pytest.raises(o2d.O2DError) as ei:
        ctx.find_vels("v2", o2d.VELANDGRAD, o2d.MEM_HOST, 4, a, a, a, a, 0, 0, None, None, None, None, None,
                      4, a, a, a, a.copy(), a.copy())
    assert ei.value.code == o2d.ERR_UNSUPPORTED
    ctx.advect("v2", 2, o2d.MEM_HOST, 0, None, None, None, None, None, None, dt=0.1)   # empty set: no-op


def test_sharded_single_rank_find_vels_and_panels(ctx, gs, port):
    """The sharded entry points with one rank (no communicator) must reproduce the plain calls."""
    g = gs
    fs = tuple(g["fs"])
    vort = (g["x"], g["y"], g["s"], g["r2"])
    pan = (g["nx"], g["ny"], g["idx"], g["pvs"], g["pss"])
    plain = ctx.find_vels_np("v2", o2d.VELANDVORT, vort, pan, g["tx"], g["ty"], g["tr"], fs)
    nt = g["tx"].size
    out = [np.empty(nt, np.float32) for _ in range(3)]
    ctx.find_vels_sharded("v2", o2d.VELANDVORT, o2d.MEM_HOST, g["x"].size, g["x"].size, *vort,
                          g["pvs"].size, g["nx"].size, g["nx"], g["ny"], g["idx"], g["pvs"], g["pss"],
                          nt, g["tx"], g["ty"], g["tr"], out[0], out[1], out[2], fs)
    for a, b in zip(out, plain):
        assert W.rel_err(a, b) < 1e-6     # same sums, grouped differently (sharded: big tiles; plain small set: fine tiles)
    area = port.panel_areas(g["nx"], g["ny"], g["idx"])
    want = ctx.points_on_panels_np(g["x"], g["y"], g["s"], g["nx"], g["ny"], g["idx"], area)
    pu = np.zeros(area.size, np.float32); pv = np.zeros(area.size, np.float32)
    ctx.points_on_panels_sharded(o2d.MEM_HOST, g["x"].size, g["x"], g["y"], g["s"], area.size, g["nx"].size,
                                 g["nx"], g["ny"], g["idx"], area, pu, pv)
    assert W.rel_err((pu, pv), want) < 1e-7


@pytest.mark.parametrize("n", [0, 1, 1000, 300001])
def test_points_status_reductions(n, ctx, port):
    """circulation / impulse / max strength.  The GPU sums in float64; the reference's impulse is a serial float
    sum whose own rounding error grows with N, so both are compared with the float64 value of the same sums."""
    import torch
    x, y, s, _ = W.gaussian_blob(max(n, 1), seed=5, signed=True)
    x, y, s = x[:n], y[:n], (s[:n] * max(n, 1)).astype(np.float32)
    got = ctx.points_status(o2d.MEM_HOST, n, x, y, s)
    want = port.points_status(x, y, s)
    s64, x64, y64 = s.astype(np.float64), x.astype(np.float64), y.astype(np.float64)
    exact = (s64.sum(), -(s64 * y64).sum(), (s64 * x64).sum(), np.abs(s64).max() if n else 1.0)
    scale = max(np.abs(s64).sum(), 1e-30)
    for k in range(3):
        assert abs(got[k] - exact[k]) <= 1e-6 * scale
        assert abs(got[k] - exact[k]) <= abs(want[k] - exact[k]) + 1e-7 * scale      # never worse than the reference
    assert got[3] == np.float32(exact[3]) == want[3]
    assert got[0] == want[0] or abs(got[0] - want[0]) <= 1e-7 * scale                 # both are float64 sums
    if n:
        d = [torch.from_numpy(a).cuda() for a in (x, y, s)]
        assert ctx.points_status(o2d.MEM_DEVICE, n, *d) == got                        # device path, bit-reproducible


def _ring_targets(cx, cy, length, nrho=240, nphi=48, lo=-1.5, hi=2.5):
    """Targets on rings around (cx, cy) at 10**lo .. 10**hi panel lengths: covers the general pair, the far-field
    series (beyond 3.16 panel lengths, panel_far_g / chorner in o2d_device.cuh) and the switch between them."""
    rho = np.logspace(lo, hi, nrho)[:, None] * length
    phi = (np.arange(nphi) + 0.37) * (2 * np.pi / nphi)
    return (cx + rho * np.cos(phi)[None, :]).astype(np.float32).ravel(), (cy + rho * np.sin(phi)[None, :]).astype(np.float32).ravel()


@pytest.mark.parametrize("length,angle", [(0.01, 0.3), (0.00123, 2.1), (0.2, -1.0)])
def test_panel_near_far_transition_pointwise(length, angle, ctx):
    """ONE panel, vortex + source strength: every target separately within 3e-6 of exact (float64) arithmetic,
    relative to the local velocity magnitude (the reference's own float32 form is up to 1e-4 off here)."""
    from oracle.f64 import panels_on_points_f64
    nx = np.array([0.3, 0.3 + length * np.cos(angle)], np.float32)
    ny = np.array([0.2, 0.2 + length * np.sin(angle)], np.float32)
    idx = np.array([0, 1], np.uint32)
    vs = np.array([0.8], np.float32); ss = np.array([-0.6], np.float32)
    tx, ty = _ring_targets(0.5 * (float(nx[0]) + float(nx[1])), 0.5 * (float(ny[0]) + float(ny[1])), length)
    u, v = ctx.panels_on_points_np(nx, ny, idx, vs, ss, tx, ty)
    eu, ev = panels_on_points_f64(nx, ny, idx, vs, ss, tx, ty)
    err = np.hypot(u - eu, v - ev) / np.hypot(eu, ev)
    assert err.max() < 3e-6, (err.max(), np.argmax(err))


def test_points_on_panels_near_far_transition_pointwise(ctx):
    """ONE particle against many panels of different length, orientation and distance: each panel's velocity is a
    single pair, so the check is pointwise.  Exact = float64."""
    from oracle.f64 import points_on_panels_f64
    rng = np.random.default_rng(11)
    npan = 6000
    length = 10 ** rng.uniform(-3, -1, npan)
    ang = rng.uniform(0, 2 * np.pi, npan)
    rho = 10 ** rng.uniform(-1.0, 2.0, npan) * length
    phi = rng.uniform(0, 2 * np.pi, npan)
    cx, cy = 0.1 + rho * np.cos(phi), -0.2 + rho * np.sin(phi)       # panel centres around the particle
    nx = np.empty(2 * npan, np.float32); ny = np.empty(2 * npan, np.float32)
    nx[0::2] = cx - 0.5 * length * np.cos(ang); nx[1::2] = cx + 0.5 * length * np.cos(ang)
    ny[0::2] = cy - 0.5 * length * np.sin(ang); ny[1::2] = cy + 0.5 * length * np.sin(ang)
    idx = np.arange(2 * npan, dtype=np.uint32)
    area = np.hypot(nx[1::2] - nx[0::2], ny[1::2] - ny[0::2]).astype(np.float32)
    sx = np.array([0.1], np.float32); sy = np.array([-0.2], np.float32); s = np.array([1.3], np.float32)
    pu, pv = ctx.points_on_panels_np(sx, sy, s, nx, ny, idx, area)
    eu, ev = points_on_panels_f64(sx, sy, s, nx, ny, idx, area)
    err = np.hypot(pu - eu, pv - ev) / np.hypot(eu, ev)
    assert err.max() < 3e-6, (err.max(), np.argmax(err))


def test_callable_from_a_non_main_thread_with_changing_sizes(ctx, port):
    """SURVEY 8b 'Threading': the GUI calls the path from a std::async worker, sequentially, with sizes that change
    between calls (diffusion / merge resize the collections)."""
    import threading
    results = {}

    def worker():
        try:
            for n in (700, 33, 2500, 1, 1024):
                sx, sy, ss, sr = W.gaussian_blob(n, seed=n)
                results[n] = (ctx.points_on_points_np("v2", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr),
                              port.points_on_points("v2", oracle.VELONLY, sx, sy, ss, sr, sx, sy, sr))
        except Exception as e:  # pragma: no cover
            results["error"] = e

    t = threading.Thread(target=worker)
    t.start(); t.join()
    assert "error" not in results, results.get("error")
    for n, (got, want) in results.items():
        assert W.rel_err(got, want[:2]) < 1e-5, n


def test_find_vels_multi_collections(ctx, port):
    """o2d_find_vels_multi: three particle collections (odd counts, one empty) and two surface collections on blob targets
    and on inert targets -- ONE call, against the reference's loop order (zero; every source in turn; finalize) done
    with the oracle; and with a single collection it is o2d_find_vels bit for bit."""
    fs = (0.7, -0.2)
    empty = tuple(np.zeros(0, np.float32) for _ in range(4))
    vorts = [W.gaussian_blob(1501, seed=41, signed=True), empty, W.gaussian_blob(2333, seed=42, signed=True),
             W.annulus_blobs(777, seed=3)]
    vorts = [tuple(np.asarray(a, np.float32) for a in v) for v in vorts]
    pans = [W.circle_panels(143, radius=0.5, seed=1), W.circle_panels(61, radius=0.2, center=(1.1, 0.3), seed=2)]
    pans = [(p[0], p[1], p[2], p[3], p[4]) for p in pans]
    pans[1] = pans[1][:4] + (None,)                       # vortex-only second surface
    for blob in (True, False):
        tx, ty, _, tr = W.gaussian_blob(1999, seed=77, sigma=0.6)
        tr = tr if blob else None
        for results in (o2d.VELONLY, o2d.VELANDVORT):
            got = ctx.find_vels_multi_np("v2", results, vorts, pans, tx, ty, tr, fs)
            u = np.zeros(tx.size, np.float32); v = np.zeros(tx.size, np.float32); w = np.zeros(tx.size, np.float32)
            for vx, vy, vs, vr in vorts:
                if vx.size == 0:
                    continue
                r = port.points_on_points("v2", results, vx, vy, vs, vr, tx, ty, tr)
                u += r[0]; v += r[1]
                if results == o2d.VELANDVORT:
                    w += r[2]
            for nx, ny, idx, pvs, pss in pans:
                pu, pv = port.panels_on_points(nx, ny, idx, pvs, pss, tx, ty)
                u += pu; v += pv
            want = port.finalize_vels(u, v, w if results == o2d.VELANDVORT else None, None, fs)
            assert W.rel_err(got[:2], want[:2]) < TOL
            if results == o2d.VELANDVORT:
                assert W.rel_err(got[2], want[2]) < TOL
    one = ctx.find_vels_multi_np("rm", o2d.VELONLY, [vorts[0]], [pans[0]], tx, ty, None, fs)
    ref = ctx.find_vels_np("rm", o2d.VELONLY, vorts[0], pans[0], tx, ty, None, fs)
    np.testing.assert_array_equal(one[0], ref[0]); np.testing.assert_array_equal(one[1], ref[1])
    # no sources at all: the freestream
    none = ctx.find_vels_multi_np("rm", o2d.VELONLY, [], [], tx, ty, None, fs)
    np.testing.assert_array_equal(none[0], np.full(tx.size, np.float32(fs[0])))


def test_sharded_entry_with_empty_block_and_exp_gradients(ctx, port):
    """A rank that owns no particles packs a filler block (zero strength, far away, small radius): it must add exactly
    nothing -- also for the EXP core's gradient results, whose pair term is NaN at zero distance, and also
    for a target at the origin."""
    sx, sy, ss, sr = W.gaussian_blob(3001, seed=8, signed=True)
    tx = np.concatenate([[0.0], sx[:500] + 0.013]).astype(np.float32)
    ty = np.concatenate([[0.0], sy[:500] - 0.007]).astype(np.float32)
    tr = np.concatenate([[sr[0]], sr[:500]]).astype(np.float32)
    for core in ("exp", "v3", "rm"):
        outs = [np.zeros(tx.size, np.float32) for _ in range(3)]
        ctx.points_on_points_sharded(core, o2d.VELANDVORT, o2d.MEM_HOST, 0, 4096, None, None, None, None,
                                     tx.size, tx, ty, tr, outs[0], outs[1], outs[2])
        for o in outs:
            np.testing.assert_array_equal(o, np.zeros_like(o))
        # one rank, block larger than the particle count: in-block padding + tile tail, same result as the plain call
        got = [np.zeros(tx.size, np.float32) for _ in range(3)]
        ctx.points_on_points_sharded(core, o2d.VELANDVORT, o2d.MEM_HOST, sx.size, 5000, sx, sy, ss, sr,
                                     tx.size, tx, ty, tr, got[0], got[1], got[2])
        want = ctx.points_on_points_np(core, o2d.VELANDVORT, sx, sy, ss, sr, tx, ty, tr)
        for a, b in zip(got, want):
            assert W.rel_err(a, b) < 1e-6


def test_find_vels_multi_device_memory_equals_host_path(ctx):
    """o2d_find_vels_multi with device pointers in the views (O2D_MEM_DEVICE) gives the bits of the host-array path."""
    import ctypes as C
    import torch
    dev = torch.device("cuda:0")
    fs = (0.3, 0.1)
    vorts = [W.gaussian_blob(3001, seed=5, signed=True), W.annulus_blobs(1200, seed=6)]
    pan = W.circle_panels(143, radius=0.5, seed=1)
    tx, ty, _, tr = W.gaussian_blob(2500, seed=7, sigma=0.7)
    host = ctx.find_vels_multi_np("v2", o2d.VELANDVORT, vorts, [pan], tx, ty, tr, fs)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    dv = [[T(a) for a in v] for v in vorts]
    dp = [T(pan[0]), T(pan[1]), T(pan[2].astype(np.int32)), T(pan[3]), T(pan[4])]
    dt = [T(tx), T(ty), T(tr)]
    out = [torch.empty(tx.size, device=dev) for _ in range(3)]
    pv = (o2d.PointsView * 2)(*[o2d.PointsView(v[0].numel(), *[t.data_ptr() for t in v]) for v in dv])
    qv = (o2d.PanelsView * 1)(o2d.PanelsView(pan[3].size, pan[0].size, *[t.data_ptr() for t in dp]))
    rc = ctx.lib.o2d_find_vels_multi(ctx._h, o2d.CORE_ID["v2"], o2d.VELANDVORT, o2d.MEM_DEVICE, 2, pv, 1, qv, tx.size,
                                     C.c_void_p(dt[0].data_ptr()), C.c_void_p(dt[1].data_ptr()), C.c_void_p(dt[2].data_ptr()),
                                     C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), C.c_void_p(out[2].data_ptr()),
                                     fs[0], fs[1])
    assert rc == o2d.OK, ctx.lib.o2d_last_error()
    ctx.synchronize()
    for a, b in zip(out, host):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
