"""Every kernel of the path at full size on the B200, next to the reference's own CPU implementation
(oracle/_ref, OpenMP on all host cores, USE_VC off) on a bounded sample of the same inputs, and the
difference between the two on that sample.  Writes a markdown table (profiles/) -- this is test/bench
infrastructure: the oracle is only the checker and the CPU arm, never the product path.

   python tools/path_table.py [--n 1000000] [--panels 20000] [--out profiles/r01_path_table.md]
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
ap.add_argument("--panels", type=int, default=20_000)
ap.add_argument("--cpu-targets", type=int, default=4096)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_path_table.md"))
args = ap.parse_args()

ctx = o2d.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
info = ctx.device_info()
cores = oracle.set_threads()
peak = info["sm_count"] * 256 * info["sm_clock_mhz"] * 1e6 / 1e12
rows = []


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def gpu_ms(fn, reps=3):
    """kernel time of the dominant pair kernel per call (library CUDA events on the launching stream)"""
    fn(); ctx.synchronize()
    ctx.reset_counters(); ctx.enable_timing(True)
    for _ in range(reps):
        fn()
    ctx.synchronize()
    c = ctx.counters(); ctx.enable_timing(False)
    return c["pair_ms"] / reps


def wall_ms(fn, reps=3):
    fn(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_s(fn):
    fn()                      # warm-up (page-in, OpenMP team)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    return best, out


def add(name, refline, units, unit_name, ms, cpu_units, cpu_sec, err, flops=None, note=""):
    g = units / (ms * 1e-3)
    c = cpu_units / cpu_sec
    frac = f"{g * flops / 1e12 / peak:.3f}" if flops else "-"
    rows.append(f"| {name} | {refline} | {units:.3g} {unit_name} | {ms:.3f} | {g:.3e} | {frac} | {c:.3e} | {g / c:.0f}x | {err:.1e} | {note} |")
    print(rows[-1], flush=True)


n = args.n
sub = W.strided_subset(n, args.cpu_targets)
FL = {"rm": (15, 28, 32, 34), "exp": (19, 35, 39, 41), "wl": (17, 31, 35, 37), "v2": (19, 32, 36, 38), "v3": (22, 35, 39, 41)}
RES = {o2d.VELONLY: 0, o2d.VELANDVORT: 1, o2d.VELANDSHEAR: 2, o2d.VELANDGRAD: 3}
RESNAME = {o2d.VELONLY: "velocity", o2d.VELANDVORT: "velocity+vorticity", o2d.VELANDSHEAR: "velocity+shear",
           o2d.VELANDGRAD: "velocity+gradient"}

# ---- particles on particles, every core / result type --------------------------------------------------------
for core, results, vary in (("rm", o2d.VELONLY, True), ("rm", o2d.VELONLY, False), ("v2", o2d.VELONLY, False),
                            ("v2", o2d.VELONLY, True), ("exp", o2d.VELONLY, True), ("wl", o2d.VELONLY, True),
                            ("v3", o2d.VELONLY, True), ("rm", o2d.VELANDVORT, True), ("rm", o2d.VELANDSHEAR, True),
                            ("rm", o2d.VELANDGRAD, True), ("v2", o2d.VELANDVORT, False)):
    x, y, s, r = W.gaussian_blob(n, seed=1, vary_radius=vary)
    d = [T(a) for a in (x, y, s, r)]
    nout = {o2d.VELONLY: 2, o2d.VELANDVORT: 3, o2d.VELANDSHEAR: 3, o2d.VELANDGRAD: 6}[results]
    outs = [torch.zeros(n, device=dev) for _ in range(nout)]

    def run():
        for o in outs:
            o.zero_()
        ctx.points_on_points(core, results, o2d.MEM_DEVICE, n, d[0], d[1], d[2], d[3], n, d[0], d[1], d[3], outs[0], outs[1],
                             outs[2] if nout == 3 else None, *(outs[2:6] if nout == 6 else ()))
    ms = gpu_ms(run)
    got = [o.cpu().numpy()[sub] for o in outs]
    ref = oracle.ref(core)
    if results == o2d.VELANDGRAD:
        sec, want = cpu_s(lambda: ref.points_on_points_grad(x, y, s, r, x[sub], y[sub], r[sub]))
    else:
        sec, want = cpu_s(lambda: ref.points_on_points(results, x, y, s, r, x[sub], y[sub], r[sub]))
    err = max(W.rel_err((got[0], got[1]), (want[0], want[1])),
              *[W.rel_err(got[k], want[k]) for k in range(2, nout)]) if nout > 2 else W.rel_err(tuple(got), tuple(want))
    add(f"points_on_points {core.upper()} {RESNAME[results]}{'' if vary else ' (uniform radius)'}",
        "`points_affect_points` Influence.h:347-403", float(n) * n, "pairs", ms, float(sub.size) * n, sec, err,
        flops=FL[core][RES[results]], note=f"CPU sample {sub.size} strided targets x all sources")
    del d, outs

# ---- Gaussian vorticity -------------------------------------------------------------------------------------------
x, y, s, r = W.gaussian_blob(n, seed=1)
d = [T(a) for a in (x, y, s, r)]
tw = torch.zeros(n, device=dev)


def run():
    tw.zero_()
    ctx.points_vorticity(o2d.MEM_DEVICE, n, d[0], d[1], d[2], d[3], n, d[0], d[1], tw)


ms = gpu_ms(run)
sec, want = cpu_s(lambda: oracle.ref("rm").points_vorticity(x, y, s, r, x[sub], y[sub]))
add("points_vorticity (Gaussian)", "`points_affect_points_vorticity` InfluenceVort.h:47-126", float(n) * n, "pairs", ms,
    float(sub.size) * n, sec, W.rel_err(tw.cpu().numpy()[sub], want), flops=13, note="as above")

# ---- panels ----------------------------------------------------------------------------------------------------------
nx, ny, idx, vs, ss = W.multi_circle_panels(4, args.panels // 4)
npan = vs.size
# smooth sheet strengths (flow-over-cylinder like), the physically relevant case: with random +-1 strengths the far
# field of 1e-4-long panels cancels and the reference's own float32 rounding noise exceeds 1e-5 of max|u|
th = 2.0 * np.pi * (np.arange(npan) % (args.panels // 4)) / (args.panels // 4)
vs = (2.0 * np.sin(th) + 0.3).astype(np.float32); ss = (0.5 * np.cos(th)).astype(np.float32)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.f64 import panels_on_points_f64, points_on_panels_f64
area = np.hypot(nx[idx[1::2]] - nx[idx[0::2]], ny[idx[1::2]] - ny[idx[0::2]]).astype(np.float32)
dp = [T(a) for a in (nx, ny, vs, ss, area)]
didx = T(idx.astype(np.int32))
tu, tv = torch.zeros(n, device=dev), torch.zeros(n, device=dev)


def run():
    tu.zero_(); tv.zero_()
    ctx.panels_on_points(o2d.MEM_DEVICE, npan, nx.size, dp[0], dp[1], didx, dp[2], dp[3], n, d[0], d[1], tu, tv)


ms = gpu_ms(run)
sec, want = cpu_s(lambda: oracle.ref("rm").panels_on_points(nx, ny, idx, vs, ss, x[sub], y[sub]))
got = (tu.cpu().numpy()[sub], tv.cpu().numpy()[sub])
s64 = sub[:512]
exact = panels_on_points_f64(nx, ny, idx, vs, ss, x[s64], y[s64])
add("panels_on_points (vortex+source)", "`panels_affect_points` Influence.h:548-602", float(npan) * n, "pairs", ms,
    float(npan) * sub.size, sec, W.rel_err(got, want), flops=47,
    note=f"{npan} panels (4 circles); CPU sample {sub.size} targets; vs float64 arithmetic on 512 targets: GPU "
         f"{W.rel_err((got[0][:512], got[1][:512]), exact):.1e}, reference {W.rel_err((want[0][:512], want[1][:512]), exact):.1e}")

pu, pv = torch.zeros(npan, device=dev), torch.zeros(npan, device=dev)


def run():
    pu.zero_(); pv.zero_()
    ctx.points_on_panels(o2d.MEM_DEVICE, n, d[0], d[1], d[2], npan, nx.size, dp[0], dp[1], didx, dp[4], pu, pv)


ms = gpu_ms(run)
psub = 256      # panels of the CPU sample: the first 256 of the first circle against all particles
sec, want = cpu_s(lambda: oracle.ref("rm").points_on_panels(x, y, s, nx, ny, idx[:2 * psub]))
got = (pu.cpu().numpy()[:psub], pv.cpu().numpy()[:psub])
exact = points_on_panels_f64(x[::8], y[::8], s[::8], nx, ny, idx[:2 * 64], area[:64])
g8 = ctx.points_on_panels_np(x[::8], y[::8], s[::8], nx, ny, idx[:2 * 64], area[:64])
r8 = oracle.ref("rm").points_on_panels(x[::8], y[::8], s[::8], nx, ny, idx[:2 * 64])
add("points_on_panels (BEM right-hand side)", "`points_affect_panels` Influence.h:719-763", float(npan) * n, "pairs", ms,
    float(psub) * n, sec, W.rel_err(got, (want[0], want[1])), flops=39,
    note=f"CPU sample {psub} panels x all particles; vs float64 arithmetic (64 panels x n/8 particles): GPU "
         f"{W.rel_err(g8, exact):.1e}, reference {W.rel_err((r8[0], r8[1]), exact):.1e}")

# ---- BEM coefficients ---------------------------------------------------------------------------------------------------
cx, cy, cidx, _, _ = W.circle_panels(args.panels, radius=0.5)
i0, i1 = cidx[0::2], cidx[1::2]
px, py = cx[i1] - cx[i0], cy[i1] - cy[i0]
carea = np.hypot(px, py).astype(np.float32)
ttx, tty = (px / carea).astype(np.float32), (py / carea).astype(np.float32)
dc = [T(a) for a in (cx, cy, carea, ttx, tty, -tty, ttx)]
dcidx = T(cidx.astype(np.int32))
npc = carea.size
cm = torch.empty((2 * npc, 2 * npc), device=dev)


def run():
    ctx.panels_on_panels_coeff(o2d.MEM_DEVICE, npc, cx.size, dc[0], dc[1], dcidx, dc[2], npc, cx.size, dc[0], dc[1], dcidx,
                               dc[2], dc[3], dc[4], dc[5], dc[6], 1, 1, 2 * npc, cm)


ms = gpu_ms(run)
nsmall = 2000
sx_, sy_, sidx_, _, _ = W.circle_panels(nsmall, radius=0.5)
sec, (cref, bases, sa) = cpu_s(lambda: oracle.ref("rm").panels_on_panels_coeff((sx_, sy_, sidx_)))
got = ctx.panels_on_panels_coeff_np((sx_, sy_, sidx_), sa, (sx_, sy_, sidx_), bases, both=True, same=True)
err = float(np.abs(got - cref).max() / np.abs(cref).max())
add("panels_on_panels_coeff (both strengths)", "`panels_on_panels_coeff` Coefficients.h:155-427", float(npc) ** 2, "pairs", ms,
    float(nsmall) ** 2, sec, err, note=f"{16 * float(npc) ** 2 / 1e9:.1f} GB matrix, store stream {16 * float(npc) ** 2 / (ms * 1e-3) / 1e9:.0f} GB/s; CPU: {nsmall} panels on themselves")
del cm

# ---- clear_inner_layer -----------------------------------------------------------------------------------------------------
snx, sny = (-py / carea).astype(np.float32), (px / carea).astype(np.float32)
rng = np.random.default_rng(1)
qx = rng.uniform(-1, 1, n).astype(np.float32); qy = rng.uniform(-1, 1, n).astype(np.float32)
dn = [T(a) for a in (snx, sny)]
state = {}


def run():
    state["x"], state["y"] = T(qx), T(qy)
    ctx.clear_inner_layer(1, o2d.MEM_DEVICE, npc, cx.size, dc[0], dc[1], dcidx, dn[0], dn[1], n, state["x"], state["y"], 0.2, 0.02)


ms = gpu_ms(run, reps=2)
csub = np.arange(0, n, n // 8192)[:8192]
sec, want = cpu_s(lambda: oracle.ref("rm").clear_inner_layer(cx, cy, cidx, qx[csub], qy[csub], 0.2, 0.02))
gx, gy = state["x"].cpu().numpy()[csub], state["y"].cpu().numpy()[csub]
err = float(max(np.abs(gx - want[0]).max(), np.abs(gy - want[1]).max()))
add("clear_inner_layer (push)", "`clear_inner_panp2` Reflect.h:490-733", float(npc) * n, "panel-point tests", ms,
    float(npc) * csub.size, sec, err, note=f"first sweep counted; error = max |dx| (absolute); CPU sample {csub.size} particles")

# ---- reflect_interior -------------------------------------------------------------------------------------------------------
def run():
    state["x"], state["y"] = T(qx), T(qy)
    ctx.reflect_interior(o2d.MEM_DEVICE, npc, cx.size, dc[0], dc[1], dcidx, dn[0], dn[1], n, state["x"], state["y"])


ms = gpu_ms(run, reps=2)
sec, want = cpu_s(lambda: oracle.ref("rm").reflect_interior(cx, cy, cidx, qx[csub], qy[csub], s[csub], r[csub]))
gx, gy = state["x"].cpu().numpy()[csub], state["y"].cpu().numpy()[csub]
err = float(max(np.abs(gx - want[0]).max(), np.abs(gy - want[1]).max()))
add("reflect_interior", "`reflect_panp2` Reflect.h:174-350", float(npc) * n, "panel-point tests", ms,
    float(npc) * csub.size, sec, err, note=f"error = max |dx| (absolute); CPU sample {csub.size} particles")

# ---- .vtu output ------------------------------------------------------------------------------------------------------------
import tempfile
u_w = np.random.default_rng(3).standard_normal((2, n)).astype(np.float32)
du = [T(a) for a in u_w]
vtu_bytes = o2d.vtu_points_size(n)
tmpd = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
vpath = os.path.join(tmpd, "part_00_00001.vtu")
kms = gpu_ms(lambda: ctx.write_vtu_points(o2d.MEM_DEVICE, vpath, n, d[0], d[1], d[2], d[3], du[0], du[1], 0.5))
t0 = time.perf_counter()
for _ in range(3):
    ctx.write_vtu_points(o2d.MEM_DEVICE, vpath, n, d[0], d[1], d[2], d[3], du[0], du[1], 0.5)
e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
got = open(vpath, "rb").read()
pinned = torch.empty(vtu_bytes, dtype=torch.uint8).pin_memory()
ctx.vtu_points(o2d.MEM_DEVICE, n, d[0], d[1], d[2], d[3], du[0], du[1], 0.5, out=pinned)
t0 = time.perf_counter()
for _ in range(5):
    ctx.vtu_points(o2d.MEM_DEVICE, n, d[0], d[1], d[2], d[3], du[0], du[1], 0.5, out=pinned)
img_ms = (time.perf_counter() - t0) / 5 * 1e3
assert pinned.numpy().tobytes() == got
t0 = time.perf_counter(); want, _ = oracle.ref("rm").write_vtk_points(tmpd, 0, 2, 0.5, x, y, s, r, u_w[0], u_w[1]); sec = time.perf_counter() - t0
nbad = 0 if got == want else 1 + sum(a != b for a, b in zip(got, want))
add("write_vtk (.vtu image of the particles)", "`Points::write_vtk` Points.h:556-740, VtkXmlWriter.h:71-106", float(vtu_bytes), "file bytes", kms,
    float(vtu_bytes), sec, float(nbad), note=f"GPU ms = the encode kernel alone ({(24 * n + vtu_bytes) / (kms * 1e-3) / 1e9:.0f} GB/s of HBM traffic); "
    f"o2d_vtu_points, finished image in pinned host memory (kernel + skeleton + D2H): {img_ms:.2f} ms = {vtu_bytes / (img_ms * 1e-3) / 1e9:.1f} GB/s; "
    f"o2d_write_vtu_points, the same plus fwrite of the image to tmpfs: {e2e_ms:.1f} ms ({sec * 1e3 / e2e_ms:.0f}x the reference's writer, whose time "
    f"also includes its file write); error column = number of differing bytes (0 = identical file)")

# ---- status reductions, zero/finalize ----------------------------------------------------------------------------------------
ms = wall_ms(lambda: ctx.points_status(o2d.MEM_DEVICE, n, d[0], d[1], d[2]), reps=5)
t0 = time.perf_counter(); want = oracle.port().points_status(x, y, s); sec = time.perf_counter() - t0
got = ctx.points_status(o2d.MEM_DEVICE, n, d[0], d[1], d[2])
err = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(got, want))
add("points_status (circulation, impulse, max strength)", "ElementBase.h:362-394, Points.h:270-285", float(n), "particles", ms,
    float(n), sec, err, note="GPU time includes the 4-float result read-back; CPU = C restatement, one thread; the reference sums serially in float32, the GPU in float64 (fixed order)")

hdr = [
    f"# Every kernel of the path at full size: B200 vs the reference's CPU implementation (round 2)",
    "",
    f"`python tools/path_table.py --n {n} --panels {args.panels}` on one B200 ({info['sm_count']} SMs, {info['sm_clock_mhz']} MHz); "
    f"device-resident inputs, time = the pair kernel's own CUDA events on the launching stream (library timers), mean of 3 launches "
    f"after one warm-up.  CPU = `oracle/_ref` (the unmodified reference templates, g++ -O3 -fopenmp, USE_VC off) on **{cores} host "
    f"cores**, best of 2 after a warm-up, on the bounded sample named in the last column; `max rel err` compares the GPU result with "
    f"that CPU result on the sample (north_star metric: max|d|/max|ref|, tolerance 1e-5).  FP32 fraction = reference flops per pair "
    f"(SURVEY 8a table) x pairs/s / {peak:.2f} TFLOP/s nominal.",
    "",
    "| entry point | replaces (reference src/) | work per call | GPU ms | GPU units/s | frac FP32 peak | CPU units/s | GPU/CPU | max rel err | notes |",
    "|---|---|---|---|---|---|---|---|---|---|",
]
with open(args.out, "w") as f:
    f.write("\n".join(hdr + rows) + "\n")
print("wrote", args.out)
