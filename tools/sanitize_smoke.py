"""Every kernel of the library once at small ragged sizes -- the target of
   compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_smoke.py
(results are also checked against the oracle so a silent corruption cannot pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import omega2d_b200 as o2d
import oracle
from omega2d_b200 import workloads as W

ctx = o2d.Context(0)
P = oracle.port()
PART = os.environ.get("O2D_SANITIZE_PART", "all")     # "pp": pair-sum kernels only, "rest": everything else, "tiny": racecheck subset
ns, nt = 2371, 1533           # not multiples of any tile
sx, sy, ss, sr = W.gaussian_blob(ns, seed=1)
tx, ty, _, tr = W.gaussian_blob(nt, seed=2)
if PART == "tiny":
    for core in ("rm", "exp"):
        for res in (o2d.VELONLY, o2d.VELANDGRAD):
            got = ctx.points_on_points_np(core, res, sx[:700], sy[:700], ss[:700], sr[:700], tx[:300], ty[:300], tr[:300])
            want = P.points_on_points_grad(core, sx[:700], sy[:700], ss[:700], sr[:700], tx[:300], ty[:300], tr[:300]) if res == o2d.VELANDGRAD \
                else P.points_on_points(core, res, sx[:700], sy[:700], ss[:700], sr[:700], tx[:300], ty[:300], tr[:300])
            assert W.rel_err(got[:2], want[:2]) < 1e-5
    # runs that cross target-block boundaries and stage rings that wrap (9 000 sources = 18 coarse tiles, 7 000 targets)
    bx, by, bs, br = W.gaussian_blob(9000, seed=5)
    got = ctx.points_on_points_np("rm", o2d.VELONLY, bx, by, bs, br, bx[:7000], by[:7000], br[:7000])
    want = P.points_on_points("rm", o2d.VELONLY, bx, by, bs, br, bx[:7000], by[:7000], br[:7000])
    assert W.rel_err(got, want) < 1e-5
    print("SANITIZE_SMOKE OK (tiny)")
    sys.exit(0)
for core in (o2d.CORES if PART in ("all", "pp") else ()):
    for res in (o2d.VELONLY, o2d.VELANDVORT, o2d.VELANDSHEAR, o2d.VELANDGRAD):
        for trr in (tr, None):
            if res == o2d.VELANDGRAD and trr is None:
                continue
            got = ctx.points_on_points_np(core, res, sx, sy, ss, sr, tx + 3.0, ty, trr)     # shifted: no coincident pairs
            want = P.points_on_points_grad(core, sx, sy, ss, sr, tx + 3.0, ty, trr) if res == o2d.VELANDGRAD else \
                P.points_on_points(core, res, sx, sy, ss, sr, tx + 3.0, ty, trr)
            assert W.rel_err(got[:2], want[:2]) < 1e-5, (core, res)
w = ctx.points_vorticity_np(sx, sy, ss, sr, tx, ty)
assert W.rel_err(w, P.points_vorticity(sx, sy, ss, sr, tx, ty)) < 1e-5
# the 256-thread / 1024-source-tile configuration (chosen for large problems), two cores, ragged count
bx, by, bs, br = W.gaussian_blob(300_007, seed=5)
sub = W.strided_subset(bx.size, 64)
for core in (("rm", "exp") if PART in ("all", "pp") else ()):
    got = ctx.points_on_points_np(core, o2d.VELANDVORT, bx, by, bs, br, bx, by, br)
    want = P.points_on_points(core, o2d.VELANDVORT, bx, by, bs, br, bx[sub], by[sub], br[sub])
    assert W.rel_err((got[0][sub], got[1][sub]), want[:2]) < 1e-5, core
nx, ny, idx, vs, sps = W.circle_panels(143)
for s2 in (sps, None):
    got = ctx.panels_on_points_np(nx, ny, idx, vs, s2, tx * 3.0, ty * 3.0)
    want = P.panels_on_points(nx, ny, idx, vs, s2, tx * 3.0, ty * 3.0)
    assert W.rel_err(got, want) < 1e-4
area = P.panel_areas(nx, ny, idx)
got = ctx.points_on_panels_np(sx * 3.0, sy * 3.0, ss, nx, ny, idx, area)
assert W.rel_err(got, P.points_on_panels(sx * 3.0, sy * 3.0, ss, nx, ny, idx, area)[:2]) < 1e-4
ttx, tty, snx, sny, a2 = P.panel_bases(nx, ny, idx)
c = ctx.panels_on_panels_coeff_np((nx, ny, idx), area, (nx, ny, idx), (ttx, tty, snx, sny, a2), both=True, same=True)
assert np.isfinite(c).all()
qx, qy = (tx * 2.5).astype(np.float32), (ty * 2.5).astype(np.float32)
ctx.clear_inner_layer_np(nx, ny, idx, snx, sny, qx, qy, 0.2, 0.02191)
ctx.reflect_interior_np(nx, ny, idx, snx, sny, qx, qy)
ctx.points_status(o2d.MEM_HOST, ns, sx, sy, ss)
img = ctx.vtu_points(o2d.MEM_HOST, nt, tx, ty, ss[:nt].copy(), sr[:nt].copy(), tx, ty, 0.5)
assert img == P.vtu_points(0.5, tx, ty, ss[:nt], sr[:nt], tx, ty)
vort = W.annulus_blobs(1777)
got = ctx.find_vels_np("v2", o2d.VELANDVORT, vort, (nx, ny, idx, vs, sps), vort[0], vort[1], vort[3], (1.0, 0.0))
want = P.find_vels("v2", o2d.VELANDVORT, vort, (nx, ny, idx, vs, sps), vort[0], vort[1], vort[3], (1.0, 0.0))
assert W.rel_err(got[:2], want[:2]) < 1e-5
for order in (1, 2, 3):
    ctx.advect_np("rm", order, *vort, 0.01, (nx, ny, idx, vs, sps), (1.0, 0.0))
# round 2: several collections in one call, the step with a body (solve callback), the sharded entry on one rank
got = ctx.find_vels_multi_np("v2", o2d.VELONLY, [vort, W.annulus_blobs(333, seed=9)], [(nx, ny, idx, vs, sps), (nx * 0.5 + 2.0, ny * 0.5, idx, vs, None)],
                             vort[0], vort[1], vort[3], (1.0, 0.0))
assert np.isfinite(got[0]).all()
near = W.annulus_blobs(1500, rin=0.495, rout=0.8, seed=4)
for order in (1, 2, 3):
    out = ctx.advect_bodies_np("v2", order, *near, (nx, ny, idx, np.zeros_like(vs), np.zeros_like(vs)), area, (snx, sny),
                               lambda stage, t, pu, pv: (-(pu * ttx + pv * tty), -(pu * snx + pv * sny)), 0.02, 0.02191, (1.0, 0.0))
    assert np.isfinite(out[0]).all() and np.isfinite(out[4]).all()
so = [np.zeros(nt, np.float32) for _ in range(3)]
ctx.points_on_points_sharded("exp", o2d.VELANDVORT, o2d.MEM_HOST, ns, ns + 5, sx, sy, ss, sr, nt, tx + 3.0, ty, tr, so[0], so[1], so[2])
assert W.rel_err(so[:2], P.points_on_points("exp", o2d.VELANDVORT, sx, sy, ss, sr, tx + 3.0, ty, tr)[:2]) < 1e-5
x, y, s, r = W.c1_speedtest(3000)
u, v = np.zeros_like(x), np.zeros_like(x)
ctx.advect("v2", 2, o2d.MEM_HOST, x.size, x.copy(), y.copy(), s, r, u, v, dt=0.005, nsteps=3)
g = o2d.Group(devices=[0, 0], min_pairs=1000)
gu, gv = np.zeros(nt, np.float32), np.zeros(nt, np.float32)
g.points_on_points("rm", o2d.VELONLY, sx, sy, ss, sr, tx, ty, tr, gu, gv)
g.close()
ctx.close()
print("SANITIZE_SMOKE OK")
