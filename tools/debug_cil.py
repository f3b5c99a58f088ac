import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import omega2d_b200 as o2d, oracle
from omega2d_b200 import workloads as W
ctx = o2d.Context(0); port = oracle.port()
npan, n = 5000, 200000
nx, ny, idx, _, _ = W.circle_panels(npan, radius=0.5)
rng = np.random.default_rng(npan)
x = rng.uniform(-0.7, 0.7, n).astype(np.float32); y = rng.uniform(-0.7, 0.7, n).astype(np.float32)
_, _, snx, sny, _ = port.panel_bases(nx, ny, idx)
want = port.clear_inner_layer(nx, ny, idx, x, y, 0.2, 0.02)
got = ctx.clear_inner_layer_np(nx, ny, idx, snx, sny, x, y, 0.2, 0.02)
d = np.hypot(got[0] - want[0], got[1] - want[1])
fl = d >= 1e-5
print("flips", fl.sum(), "moved", got[2], want[2])
r0 = np.hypot(x, y)
print("orig radius of flips: min/median/max", r0[fl].min(), np.median(r0[fl]), r0[fl].max())
print("d of flips: min/median/max", d[fl].min(), np.median(d[fl]), d[fl].max())
rg, rw = np.hypot(got[0], got[1]), np.hypot(want[0], want[1])
print("radius got vs want (flips):", np.abs(rg - rw)[fl].max())
k = np.where(fl)[0][:5]
for i in k:
    print(i, x[i], y[i], r0[i], "got", got[0][i], got[1][i], rg[i], "want", want[0][i], want[1][i], rw[i])
