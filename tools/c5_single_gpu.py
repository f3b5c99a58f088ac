"""BASELINE config C5 at full size on ONE B200: 16 777 216 particles, Winckelmans-Leonard core, velocity -- one
evaluation (2.8e14 pairs), checked against the reference's CPU loop on strided targets.  ~100 s of GPU time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import omega2d_b200 as o2d, oracle
from omega2d_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
x, y, s, r = W.gaussian_blob(n, seed=2024)
d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
u = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
ctx.reset_counters(); ctx.enable_timing(True)
t0 = time.perf_counter()
ctx.points_on_points("wl", o2d.VELONLY, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], u, v)
ctx.synchronize()
wall = time.perf_counter() - t0
ms = ctx.counters()["pair_ms"]
sub = W.strided_subset(n, 256)
oracle.set_threads()
want = oracle.ref("wl").points_on_points(oracle.VELONLY, x, y, s, r, x[sub], y[sub], r[sub])
hu, hv = u.cpu().numpy(), v.cpu().numpy()
scale = np.hypot(hu.astype(np.float64), hv.astype(np.float64)).max()
err = np.hypot(hu[sub] - want[0], hv[sub] - want[1]).max() / scale
print(f"C5 single GPU: n={n} pairs={float(n)*n:.3e} kernel {ms:.1f} ms (wall {wall*1e3:.1f}) {float(n)*n/(ms*1e-3):.4e} pairs/s "
      f"= {17*float(n)*n/(ms*1e-3)/74.45e12*100:.1f} % of FP32 peak; max rel err vs reference on 256 strided targets {err:.2e}", flush=True)
assert err < 1e-5 and np.isfinite(hu).all() and np.isfinite(hv).all()
