"""Profiling driver: runs ONE operation of the path twice (warm-up + the launch ncu captures) at bench size.
    python tools/ncu_targets.py pp:rm:2:0 | pp:<core>:<results>:<uniform> | panels | ptpan | coeff | cil | reflect   [n] [npanels]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
npan = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
if what.startswith("pp:"):
    _, core, res, uni = what.split(":")
    res, uni = int(res), int(uni)
    x, y, s, r = W.gaussian_blob(n, seed=2024, vary_radius=not uni)
    d = [T(a) for a in (x, y, s, r)]
    nout = {1: 2, 2: 6, 4: 3, 5: 3}[res]
    o = [torch.zeros(n, device=dev) for _ in range(nout)]
    tq = o[2] if nout == 3 else None
    g = o[2:6] if nout == 6 else [None] * 4
    fn = lambda: ctx.points_on_points(core, res, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], o[0], o[1], tq, *g)
else:
    x, y, s, r = W.gaussian_blob(n, seed=2024)
    nx, ny, idx, vs, ss = W.multi_circle_panels(4, npan // 4)
    import oracle
    ttx, tty, snx, sny, area = oracle.port().panel_bases(nx, ny, idx)
    dp = [T(a) for a in (nx, ny, idx.astype(np.int32), vs, ss, area, snx, sny)]
    dx, dy, ds = T(x), T(y), T(s)
    if what == "panels":
        u = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
        fn = lambda: ctx.panels_on_points(o2d.MEM_DEVICE, vs.size, nx.size, dp[0], dp[1], dp[2], dp[3], dp[4], n, dx, dy, u, v)
    elif what == "ptpan":
        pu = torch.zeros(vs.size, device=dev); pv = torch.zeros(vs.size, device=dev)
        fn = lambda: ctx.points_on_panels(o2d.MEM_DEVICE, n, dx, dy, ds, vs.size, nx.size, dp[0], dp[1], dp[2], dp[5], pu, pv)
    elif what == "coeff":
        # BEM block of one 20 000-panel circle on itself, both strengths unknown (the panel_bench.py case)
        cnx, cny, cidx, _, _ = W.circle_panels(npan, radius=0.5)
        i0, i1 = cidx[0::2], cidx[1::2]
        px, py = cnx[i1] - cnx[i0], cny[i1] - cny[i0]
        car = np.hypot(px, py).astype(np.float32)
        c = [T(np.ascontiguousarray(a)) for a in (cnx, cny, car, (px / car).astype(np.float32), (py / car).astype(np.float32),
                                                  (-py / car).astype(np.float32), (px / car).astype(np.float32))]
        ci = T(cidx.astype(np.int32))
        m = torch.empty((2 * npan, 2 * npan), device=dev)
        fn = lambda: ctx.panels_on_panels_coeff(o2d.MEM_DEVICE, npan, npan, c[0], c[1], ci, c[2], npan, npan, c[0], c[1], ci,
                                                c[2], c[3], c[4], c[5], c[6], 1, 1, 2 * npan, m)
    else:
        rng = np.random.default_rng(1)
        qx = T(rng.uniform(-1, 1, n).astype(np.float32)); qy = T(rng.uniform(0.5, 2.5, n).astype(np.float32))
        if what == "cil":
            fn = lambda: ctx.clear_inner_layer(1, o2d.MEM_DEVICE, vs.size, nx.size, dp[0], dp[1], dp[2], dp[6], dp[7], n, qx.clone(), qy.clone(), 0.2, 0.02)
        else:
            fn = lambda: ctx.reflect_interior(o2d.MEM_DEVICE, vs.size, nx.size, dp[0], dp[1], dp[2], dp[6], dp[7], n, qx.clone(), qy.clone())
fn(); ctx.synchronize()
fn(); ctx.synchronize()
print("done", what)
