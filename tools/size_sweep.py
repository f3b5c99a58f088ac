"""Development tool: pair-sum throughput against particle count (device-resident, library event timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
clk = 148 * 1.965e9
for core, uniform in (("rm", False), ("v2", True)):
    for n in (1000, 5000, 20000, 50000, 100000, 200000, 400000, 700000, 1000000):
        x, y, s, r = W.gaussian_blob(n, seed=3, vary_radius=not uniform)
        d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
        u = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
        fn = lambda: ctx.points_on_points(core, o2d.VELONLY, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], u, v)
        fn(); ctx.synchronize()
        reps = 20 if n <= 100000 else 3
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        ctx.reset_counters(); ctx.enable_timing(True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream); torch.cuda.synchronize()
        c = ctx.counters(); ctx.enable_timing(False)
        ms_call = e0.elapsed_time(e1) / reps; ms_k = c["pair_ms"] / reps
        print(f"{core} uniform={int(uniform)} n={n:8d}  call {ms_call:9.4f} ms  pair kernel {ms_k:9.4f} ms  "
              f"{n*float(n)/(ms_call*1e-3):.3e} pairs/s  {n*float(n)/(ms_k*1e-3)/clk:.2f} pairs/clk/SM (kernel)", flush=True)
