"""Development tool (debug build only): distribution of per-CTA run times of the pair-sum launch.
    O2D_LIB=build/dbg/libomega2d_b200_dbg.so python tools/cta_times.py [n] [core]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
core = sys.argv[2] if len(sys.argv) > 2 else "rm"
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
x, y, s, r = W.gaussian_blob(n, seed=3)
d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
u = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
for _ in range(2):
    ctx.points_on_points(core, 1, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], u, v)
ctx.synchronize()
lib = o2d.load_library()
N = 16384
buf = (C.c_ulonglong * (3 * N))()
lib.o2d_debug_cta_times(buf, N)
a = np.frombuffer(buf, dtype=np.uint64).reshape(N, 3).astype(np.int64)
a = a[a[:, 2] > 0]
t0 = a[:, 1].min()
dur = (a[:, 2] - a[:, 1]) * 1e-6
start = (a[:, 1] - t0) * 1e-6; end = (a[:, 2] - t0) * 1e-6
print(f"{len(a)} CTAs; start max {start.max():.3f} ms; duration ms: min {dur.min():.3f} p10 {np.percentile(dur,10):.3f} "
      f"median {np.median(dur):.3f} p90 {np.percentile(dur,90):.3f} max {dur.max():.3f}; last end {end.max():.3f} ms")
sm = a[:, 0]
per_sm = {}
for k in range(len(a)):
    per_sm.setdefault(int(sm[k]), []).append(dur[k])
keys = np.array(sorted(per_sm))
means = np.array([np.mean(per_sm[k]) for k in keys]); cnt = np.array([len(per_sm[k]) for k in keys])
print(f"{len(per_sm)} SMs; CTAs per SM min {cnt.min()} max {cnt.max()}; per-SM mean duration min {means.min():.3f} max {means.max():.3f}")
print("per-SM mean duration by smid:", " ".join(f"{keys[i]}:{means[i]:.1f}" for i in range(len(keys))))
