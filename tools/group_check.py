"""One host process, all GPUs of the box (o2d_group_*): parity on a strided target sample against the reference's CPU
loop, bit-equality with the single-device result, and wall-clock throughput with HOST arrays for 1, 2, ... N devices.

   python tools/group_check.py [--n 1000000] [--out profiles/r01_group.md]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import omega2d_b200 as o2d
import oracle
from omega2d_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--out", default=None)
args = ap.parse_args()
n = args.n
ndev = torch.cuda.device_count()
oracle.set_threads()
sx, sy, ss, sr = W.gaussian_blob(n, seed=2024)
panels = W.multi_circle_panels(4, 5000)
sub = W.strided_subset(n, 2048)
ref = oracle.ref("rm") if oracle.have_ref() else None
want = (ref.points_on_points(oracle.VELONLY, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub]) if ref
        else oracle.port().points_on_points("rm", oracle.VELONLY, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub]))
lines = []
base = None
first = None
counts = sorted({1, 2, 4, ndev} & set(range(1, ndev + 1)))
for k in counts:
    g = o2d.Group(devices=list(range(k)))
    u, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    g.points_on_points("rm", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr, u, v)      # warm-up: buffers, clocks
    best = 1e30
    for _ in range(3):
        u[:] = 0; v[:] = 0
        t0 = time.perf_counter()
        g.points_on_points("rm", o2d.VELONLY, sx, sy, ss, sr, sx, sy, sr, u, v)
        best = min(best, time.perf_counter() - t0)
    err = W.rel_err((u[sub], v[sub]), want[:2])
    if first is None:
        first = (u.copy(), v.copy())
    same = bool(np.array_equal(u, first[0]) and np.array_equal(v, first[1]))
    rate = float(n) * n / best
    base = base or rate
    # find_vels with the C4 panel set on the same particles
    fu, fv = np.empty(n, np.float32), np.empty(n, np.float32)
    g.find_vels("rm", o2d.VELONLY, (sx, sy, ss, sr), panels, sx, sy, sr, fu, fv)
    t0 = time.perf_counter()
    g.find_vels("rm", o2d.VELONLY, (sx, sy, ss, sr), panels, sx, sy, sr, fu, fv)
    fv_ms = (time.perf_counter() - t0) * 1e3
    lines.append(f"| {k} | {best * 1e3:.1f} | {rate:.3e} | {rate / base / k * 100:.1f} % | {err:.1e} | {same} | {fv_ms:.1f} |")
    print(lines[-1], flush=True)
    assert err < 1e-5
    g.close()
hdr = [f"# One process, several GPUs (o2d_group_points_on_points / o2d_group_find_vels), {n} particles, RM core, HOST arrays",
       "",
       f"`python tools/group_check.py --n {n}` on a {ndev}-GPU B200 box: ONE host process, one host thread per device, the target",
       "collection cut into contiguous blocks, every device uploads all sources over its own PCIe link.  Wall clock around the call",
       "(upload + kernels + download), best of 3.  `max rel err` = GPU vs the reference's OpenMP loop on 2048 strided targets;",
       "`same bits` = the full result equals the 1-device result bit for bit.  find_vels adds 20 000 panels (4 circles) and the finalize.",
       "",
       "| devices | ms per evaluation | interactions/s | efficiency vs 1 device | max rel err | same bits | find_vels ms |",
       "|---|---|---|---|---|---|---|"]
text = "\n".join(hdr + lines) + "\n"
if args.out:
    open(args.out, "w").write(text)
print(text)
print("GROUP_CHECK PASS")
