"""BASELINE config C2 (examples/flow_over_circle.json stand-in): one Runge-Kutta substep's worth of the
path around a 143-panel unit-diameter circle -- find_vels on the particles (particles + vortex/source panels,
Vatistas-2 core, freestream (1,0)), the BEM right-hand side (particles on panel centres), clear_inner_layer after the
move, and the status reductions -- through the C ABI with HOST arrays (what the host mirror passes), next to the
reference's own loops (oracle/_ref, OpenMP on all host cores) on the SAME full inputs, with the difference between them.

   python tools/c2_substep.py [--out profiles/r01_c2_substep.md]
"""
import argparse
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import omega2d_b200 as o2d
import oracle
from omega2d_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_c2_substep.md"))
ap.add_argument("--sizes", default="10000,100000")
args = ap.parse_args()

ctx = o2d.Context(0)
R = oracle.ref("v2")
cores = oracle.set_threads()
NP = 143
IPS = 0.02191
CF = 0.5 / math.sqrt(2.0 * math.pi)
nx, ny, idx, pvs, pss = W.circle_panels(NP, radius=0.5)
ttx, tty, snx, sny, area = oracle.port().panel_bases(nx, ny, idx)
fs = (1.0, 0.0)
rows = []


def best(fn, reps=5):
    fn()
    t = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); t = min(t, time.perf_counter() - t0)
    return t * 1e3, out


for n in [int(s) for s in args.sizes.split(",")]:
    x, y, s, r = W.annulus_blobs(n)
    # find_vels on the particles
    gms, g = best(lambda: ctx.find_vels_np("v2", o2d.VELONLY, (x, y, s, r), (nx, ny, idx, pvs, pss), x, y, r, fs))
    cms, c = best(lambda: R.find_vels(oracle.VELONLY, (x, y, s, r), (nx, ny, idx, pvs, pss), x, y, r, fs), reps=1)
    rows.append((n, "find_vels (particles + 143 panels on the particles, finalize)", float(n) * (n + NP), gms, cms,
                 W.rel_err(g[:2], c[:2])))
    # BEM right-hand side: particles on panel centres
    gms, g = best(lambda: ctx.points_on_panels_np(x, y, s, nx, ny, idx, area))
    cms, c = best(lambda: R.points_on_panels(x, y, s, nx, ny, idx), reps=2)
    rows.append((n, "points_affect_panels (BEM right-hand side)", float(n) * NP, gms, cms, W.rel_err(g, c[:2])))
    # clear_inner_layer after a move: pull a tenth of the particles into the layer first
    qx, qy = x.copy(), y.copy()
    qx[::10] *= 0.5 / np.hypot(x[::10], y[::10]) * 1.004
    qy[::10] *= 0.5 / np.hypot(x[::10], y[::10]) * 1.004
    gms, g = best(lambda: ctx.clear_inner_layer_np(nx, ny, idx, snx, sny, qx, qy, CF, IPS))
    cms, c = best(lambda: R.clear_inner_layer(nx, ny, idx, qx, qy, CF, IPS), reps=2)
    rows.append((n, f"clear_inner_layer (push; {g[2]} pushes)", float(n) * NP, gms, cms,
                 float(max(np.abs(g[0] - c[0]).max(), np.abs(g[1] - c[1]).max()))))
    # one whole particle-only RK2 step for scale (the BEM solve between the stages is the reference's Eigen GMRES)
    gms, g = best(lambda: ctx.advect_np("v2", 2, x, y, s, r, 0.01, (nx, ny, idx, pvs, pss), fs))
    cms, c = best(lambda: R.advect(2, x, y, s, r, 0.01, (nx, ny, idx, pvs, pss), fs), reps=1)
    rows.append((n, "advect, Ralston RK2 with the panel strengths held fixed (2 find_vels + moves)", 2.0 * n * (n + NP), gms, cms,
                 float(max(np.abs(g[0] - c[0]).max(), np.abs(g[1] - c[1]).max()))))

    # the reference's whole RK2 step WITH the body: per stage BEM right-hand side -> solve -> find_vels -> move ->
    # clear_inner_layer, particles resident on the GPU (o2d_advect_bodies); the linear solve is a stand-in callback
    # (strengths cancelling the slip at the panel centres), identical on both sides and outside the path
    def solve(stage, time_, pu, pv):
        return (-(pu * ttx + pv * tty)).astype(np.float32), (-(pu * snx + pv * sny)).astype(np.float32)
    z = np.zeros(NP, np.float32)
    gms, g = best(lambda: ctx.advect_bodies_np("v2", 2, qx, qy, s, r, (nx, ny, idx, z, z), area, (snx, sny), solve, 0.01, IPS, fs))
    cms, c = best(lambda: R.advect_bodies(2, qx, qy, s, r, (nx, ny, idx, z, z), solve, 0.01, IPS, fs), reps=1)
    d = np.hypot(g[0] - c[0], g[1] - c[1])
    rows.append((n, "advect_bodies, Ralston RK2 with the body: 2 x (BEM RHS, solve callback, find_vels, move, clear_inner_layer), "
                    "ONE upload + ONE download of the particles" + (f" [{int((d > 2e-6).sum())} push-threshold flips]" if (d > 2e-6).any() else ""),
                 2.0 * n * (n + 2 * NP), gms, cms, float(np.median(d)) if (d > 2e-6).any() else float(d.max())))

info = ctx.device_info()
with open(args.out, "w") as f:
    f.write(f"# C2 (flow_over_circle stand-in): one substep of the path around a {NP}-panel circle, B200 vs the reference's CPU loops\n\n")
    f.write(f"`python tools/c2_substep.py` on one B200 ({info['sm_count']} SMs); every call goes through the C ABI with HOST numpy arrays "
            f"(upload, kernels, download inside the timed region, best of 5 wall-clock); CPU = `oracle/_ref` (unmodified reference "
            f"templates, OpenMP, USE_VC off) on **{cores} host cores** on the same full inputs; error = max|d|/max|ref| for velocities "
            f"(tolerance 1e-5), max |dx| for positions.  Blobs in the annulus 0.52 < rho < 1.5, r = vdelta = 0.043818, V2 core, Uinf = (1,0).\n\n")
    f.write("| particles | operation | pairs | B200 ms (host arrays) | CPU ms | speed-up | error |\n|---|---|---|---|---|---|---|\n")
    for n, what, pairs, gms, cms, err in rows:
        f.write(f"| {n} | {what} | {pairs:.3g} | {gms:.3f} | {cms:.1f} | {cms / gms:.0f}x | {err:.1e} |\n")
print(open(args.out).read())
