"""Development tool: pair-sum kernel rate of every core at one particle count (device-resident, library event timing).
    python tools/core_sweep.py [n] [reps]      O2D_LIB=<other build> for A/B"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
info = ctx.device_info()
clk = info["sm_count"] * info["sm_clock_mhz"] * 1e6
peak = info["fp32_peak_tflops"]
print(f"lib={o2d.LIB_PATH} n={n}")
cases = [(c, o2d.VELONLY, u) for c in ("rm", "v2", "wl", "exp", "v3") for u in (False, True)]
cases += [("rm", o2d.VELANDGRAD, False), ("rm", o2d.VELANDVORT, False), ("v2", o2d.VELANDVORT, True), ("rm", o2d.VELANDSHEAR, False)]
for core, res, uniform in cases:
    x, y, s, r = W.gaussian_blob(n, seed=3, vary_radius=not uniform)
    d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
    nout = {1: 2, 2: 6, 4: 3, 5: 3}[res]
    o = [torch.zeros(n, device=dev) for _ in range(nout)]
    tq = o[2] if nout == 3 else None
    g = o[2:6] if nout == 6 else [None] * 4
    fn = lambda: ctx.points_on_points(core, res, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], o[0], o[1], tq, *g)
    fn(); ctx.synchronize()
    ctx.reset_counters(); ctx.enable_timing(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); torch.cuda.synchronize()
    c = ctx.counters(); ctx.enable_timing(False)
    ms_k = c["pair_ms"] / reps; ms_call = e0.elapsed_time(e1) / reps
    rate = n * float(n) / (ms_k * 1e-3)
    fpp = o2d.flops_per_pair(0, core, res, True)
    print(f"{core:3s} res={res} uniform={int(uniform)}  call {ms_call:9.3f} ms  kernel {ms_k:9.3f} ms  {rate:.3e} pairs/s  "
          f"{rate/clk:6.2f} pairs/clk/SM  {fpp} fl -> {fpp*rate/1e12/peak:.3f} of FP32 peak", flush=True)
