"""Development tool: pair-sum rate of one core/result at n particles, with the error against the reference CPU loop.
   python tools/core_rate.py v3 1 1000000 [uniform]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import omega2d_b200 as o2d, oracle
from omega2d_b200 import workloads as W
core = sys.argv[1]; results = int(sys.argv[2]); n = int(sys.argv[3]); uniform = len(sys.argv) > 4
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
x, y, s, r = W.gaussian_blob(n, seed=3, vary_radius=not uniform)
d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
nout = {1: 2, 2: 6, 4: 3, 5: 3}[results]
outs = [torch.zeros(n, device=dev) for _ in range(nout)]
def fn():
    for o in outs: o.zero_()
    ctx.points_on_points(core, results, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], outs[0], outs[1],
                         outs[2] if nout == 3 else None, *(outs[2:6] if nout == 6 else [None] * 4))
fn(); ctx.synchronize()
ctx.reset_counters(); ctx.enable_timing(True)
for _ in range(3): fn()
ctx.synchronize(); c = ctx.counters(); ctx.enable_timing(False)
ms = c["pair_ms"] / 3
sub = W.strided_subset(n, 2048)
oracle.set_threads()
res_cpu = 1 if results == 2 else results
want = oracle.ref(core).points_on_points(res_cpu, x, y, s, r, x[sub], y[sub], r[sub])
got = [o.cpu().numpy()[sub] for o in outs]
err = W.rel_err((got[0], got[1]), want[:2])
errq = W.rel_err(got[2], want[2]) if nout == 3 else float("nan")
print(f"{core} res{results} n={n} uniform={int(uniform)}: {ms:.3f} ms  {n*float(n)/(ms*1e-3):.4e} pairs/s  "
      f"{n*float(n)/(ms*1e-3)/(148*1.965e9):.2f} pairs/clk/SM  err_uv={err:.2e} err_q={errq:.2e}", flush=True)
