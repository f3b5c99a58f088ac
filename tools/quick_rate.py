"""Development tool: kernel rate of a few (core, results, uniform) cases; env knobs are read by the library.
    python tools/quick_rate.py n reps case[,case...]   case = core:res:uniform  e.g. rm:1:0"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
n = int(sys.argv[1]); reps = int(sys.argv[2])
cases = [c.split(":") for c in sys.argv[3].split(",")]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("O2D_"))
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
info = ctx.device_info()
clk = info["sm_count"] * info["sm_clock_mhz"] * 1e6
for core, res, uniform in cases:
    res = int(res); uniform = int(uniform)
    x, y, s, r = W.gaussian_blob(n, seed=3, vary_radius=not uniform)
    d = [torch.from_numpy(a).to(dev) for a in (x, y, s, r)]
    nout = {1: 2, 2: 6, 4: 3, 5: 3}[res]
    o = [torch.zeros(n, device=dev) for _ in range(nout)]
    tq = o[2] if nout == 3 else None
    g = o[2:6] if nout == 6 else [None] * 4
    fn = lambda: ctx.points_on_points(core, res, o2d.MEM_DEVICE, n, *d, n, d[0], d[1], d[3], o[0], o[1], tq, *g)
    fn(); ctx.synchronize()
    ctx.reset_counters(); ctx.enable_timing(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    c = ctx.counters(); ctx.enable_timing(False)
    ms_k = c["pair_ms"] / reps
    rate = n * float(n) / (ms_k * 1e-3)
    print(f"[{tag}] {core} res={res} uni={uniform} n={n} kernel {ms_k:9.3f} ms {rate:.3e} pairs/s {rate/clk:6.2f} pairs/clk/SM chk={float(o[0].double().sum()):.9e}", flush=True)
