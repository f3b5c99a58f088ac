"""Development tool: throughput of the panel kernels (device-resident arrays, library event timing).
   python tools/panel_bench.py [path/to/libomega2d_b200.so]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import omega2d_b200 as o2d
from omega2d_b200 import workloads as W
if len(sys.argv) > 1:
    o2d.LIB_PATH = os.path.abspath(sys.argv[1])
ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
clk = 148 * 1.965e9
for npan, nt in ((20000, 1 << 20), (143, 1 << 20), (5000, 1 << 18)):
    nx, ny, idx, vs, ss = W.multi_circle_panels(4, npan // 4) if npan >= 1000 else W.circle_panels(npan)
    npan = vs.size
    x, y, s, r = W.gaussian_blob(nt, seed=3)
    d = [torch.from_numpy(a).to(dev) for a in (nx, ny, vs, ss, x, y, s)]
    didx = torch.from_numpy(idx.astype(np.int32)).to(dev)
    tu = torch.zeros(nt, device=dev); tv = torch.zeros(nt, device=dev)
    area = torch.from_numpy(np.hypot(nx[idx[1::2]] - nx[idx[0::2]], ny[idx[1::2]] - ny[idx[0::2]]).astype(np.float32)).to(dev)
    pu = torch.zeros(npan, device=dev); pv = torch.zeros(npan, device=dev)
    for name, fn in (("panels_on_points", lambda: ctx.panels_on_points(o2d.MEM_DEVICE, npan, nx.size, d[0], d[1], didx, d[2], d[3], nt, d[4], d[5], tu, tv)),
                     ("points_on_panels", lambda: ctx.points_on_panels(o2d.MEM_DEVICE, nt, d[4], d[5], d[6], npan, nx.size, d[0], d[1], didx, area, pu, pv))):
        fn(); ctx.synchronize()
        ctx.reset_counters(); ctx.enable_timing(True)
        for _ in range(3):
            fn()
        ctx.synchronize()
        c = ctx.counters(); ctx.enable_timing(False)
        ms = c["pair_ms"] / 3
        rate = npan * nt / (ms * 1e-3)
        print(f"{name:18s} np={npan:6d} nt={nt:8d}  {ms:9.3f} ms  {rate:.3e} pairs/s  {rate/clk:.2f} pairs/clk/SM", flush=True)

# BEM coefficient assembly (device memory), both strengths unknown
for npan in (143, 5000, 20000):
    nx, ny, idx, _, _ = W.circle_panels(npan, radius=0.5)
    i0, i1 = idx[0::2], idx[1::2]
    px, py = nx[i1] - nx[i0], ny[i1] - ny[i0]
    area = np.hypot(px, py).astype(np.float32)
    ttx, tty = (px / area).astype(np.float32), (py / area).astype(np.float32)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (nx, ny, area, ttx, tty, -tty, ttx)]
    didx = torch.from_numpy(idx.astype(np.int32)).to(dev)
    c = torch.empty((2 * npan, 2 * npan), device=dev)
    fn = lambda: ctx.panels_on_panels_coeff(o2d.MEM_DEVICE, npan, npan, d[0], d[1], didx, d[2], npan, npan, d[0], d[1], didx,
                                            d[2], d[3], d[4], d[5], d[6], 1, 1, 2 * npan, c)
    fn(); ctx.synchronize()
    ctx.reset_counters(); ctx.enable_timing(True)
    for _ in range(3):
        fn()
    ctx.synchronize()
    cnt = ctx.counters(); ctx.enable_timing(False)
    ms = cnt["pair_ms"] / 3
    pairs = float(npan) ** 2
    print(f"panels_on_panels_coeff np={npan:6d}  {ms:9.3f} ms  {pairs/(ms*1e-3):.3e} pairs/s  matrix {16*pairs/1e9:.2f} GB  "
          f"store {16*pairs/(ms*1e-3)/1e9:.0f} GB/s", flush=True)
    del c

# clear_inner_layer (push), device memory
for npan, n in ((143, 1 << 20), (20000, 1 << 20)):
    nx, ny, idx, _, _ = W.circle_panels(npan, radius=0.5)
    i0, i1 = idx[0::2], idx[1::2]
    px, py = nx[i1] - nx[i0], ny[i1] - ny[i0]
    area = np.hypot(px, py)
    snx, sny = (-py / area).astype(np.float32), (px / area).astype(np.float32)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, n).astype(np.float32); y = rng.uniform(-1, 1, n).astype(np.float32)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (nx, ny, snx, sny)]
    didx = torch.from_numpy(idx.astype(np.int32)).to(dev)
    for rep in range(2):
        dx = torch.from_numpy(x).to(dev); dy = torch.from_numpy(y).to(dev)
        ctx.synchronize(); ctx.reset_counters(); ctx.enable_timing(True)
        moved = ctx.clear_inner_layer(1, o2d.MEM_DEVICE, npan, npan, d[0], d[1], didx, d[2], d[3], n, dx, dy, 0.2, 0.02)
        ctx.synchronize(); cnt = ctx.counters(); ctx.enable_timing(False)
    ms = cnt["pair_ms"]
    print(f"clear_inner_layer np={npan:6d} n={n:8d}  {ms:9.3f} ms  moved {moved}  {npan*float(n)/(ms*1e-3):.3e} panel-point tests/s (first sweep)", flush=True)
