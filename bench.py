#!/usr/bin/env python
"""bench.py -- Biot-Savart pair interactions per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c3vel|c1|c4|c5|v2]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one evaluation of the hot path over one synthetic particle set: the
points-on-points direct sum that Convection::find_vels runs per Runge-Kutta substep
(reference src/Convection.h:172-223 -> src/Influence.h:347-361).

Default workload (N=1): BASELINE.json configs[2], "C3": 1M-particle Gaussian vortex blob,
Rosenhead-Moore core, per-particle radii, velocity + gradient on blob targets (kernelug_0v_0b): 1e12 pairs/step.
The velocity-only evaluation of the same particles (kernelu_0v_0b, what find_vels runs per substep) is reported
under `variants` (and is `--workload c3vel`).
At N>1 GPUs the particle count follows the weak-scaling recipe of SURVEY 8d/C5 (N_P = N_1 sqrt(P),
constant pairs per GPU); targets are sharded in contiguous blocks and every rank's packed source
block is all-gathered once per step over NVLink (o2d_points_on_points_sharded).

Printed keys (one JSON line from rank 0):
  value     whole-job interactions/s, inputs resident in HBM, CUDA-event time, max over ranks
  e2e       same metric through the C ABI with HOST (pinned) buffers: H2D of sources/targets/outputs and
            D2H of the velocities inside the timed region
  roofline  dominant pair-sum kernel: reference-counted flops per pair x pairs / kernel time (events inside
            the library, on the launching stream) against the FP32 non-tensor peak
  cpu_baseline  the reference's own OpenMP loop (oracle/_ref) on this box's host cores, bounded target sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG=VERSION|INFO) go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

from omega2d_b200 import sharding, workloads as W  # noqa: E402

METRIC = "biot_savart_pair_interactions_per_sec"
UNIT = "interactions/s"

WORKLOADS = {
    # name: (generator, core, results, description)
    "c3": dict(n=1_000_000, core="rm", results=2, blob=True,
               desc="C3: synthetic 1M-particle Gaussian vortex blob, Rosenhead-Moore core, per-particle radii, "
                    "velocity + gradient on blob targets (kernelug_0v_0b: u, v and the 2x2 velocity gradient)"),
    "c3vel": dict(n=1_000_000, core="rm", results=1, blob=True,
                  desc="C3 particles, velocity only on blob targets (kernelu_0v_0b: what find_vels evaluates every substep)"),
    "c1": dict(n=20_000, core="v2", results=1, blob=True,
               desc="C1: speedtest20kx10 restated, 20k blobs, Vatistas-2 core, velocity"),
    "c4": dict(n=4_194_304, core="rm", results=1, blob=True, panels=(4, 5000), scaling="strong",
               desc="C4: 4M-particle Gaussian blob + 20k panels (4 closed circles of 5000), Rosenhead-Moore core: one "
                    "find_vels on the particles (zero, particles-on-particles, panels-on-particles, finalize)"),
    "c5": dict(n=16_777_216, core="wl", results=1, blob=True,
               desc="C5: 16M-particle Gaussian blob, Winckelmans-Leonard core, velocity"),
    "v2": dict(n=1_000_000, core="v2", results=1, blob=True,
               desc="1M-particle Gaussian vortex blob, Vatistas-2 core (reference default), velocity"),
}


def make_particles(name: str, n: int, vary_radius: bool = True):
    if name == "c1":
        return W.c1_speedtest(n)
    return W.gaussian_blob(n, seed=2024, vary_radius=vary_radius)


def ncu_traffic(core: str, results: int, n: int):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); None when no capture
    of this exact (core, results, particle count) is on file."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return None
    e = tab.get(f"{core}/res{results}/n{n}")
    return None if e is None else e


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--particles", dest="n", type=int, default=0, help="override the particle count at 1 GPU")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="default: the workload's own (weak, except c4: strong)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of one cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-size C4 block of the default (c3) run")
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--strong-particles", type=int, default=0, help="particle count of the C4 block (default: C4's 4 194 304; tests shrink it)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU side: the reference's own loop (oracle/_ref) or, if that was never built, the C port
# ---------------------------------------------------------------------------------------------
def cpu_checker(core: str, panels=None):
    """The CPU arm: the reference's own loops (oracle/_ref) or, if that was never built, the C port.  With
    panels the call is the whole find_vels (zero, particles, panels, finalize), else points_affect_points."""
    import oracle
    if oracle.have_ref():
        r = oracle.ref(core)
        if panels is not None:
            return "reference", (lambda res, sx, sy, ss, sr, tx, ty, tr: r.find_vels(res, (sx, sy, ss, sr), panels, tx, ty, tr))
        return "reference", (lambda res, *a: r.points_on_points_grad(*a) if res == 2 else r.points_on_points(res, *a))
    p = oracle.port()
    if panels is not None:
        return "port", (lambda res, sx, sy, ss, sr, tx, ty, tr: p.find_vels(core, res, (sx, sy, ss, sr), panels, tx, ty, tr))
    return "port", (lambda res, *a: p.points_on_points_grad(core, *a) if res == 2 else p.points_on_points(core, res, *a))


def cpu_sample(wl, parts, seconds: float):
    """Time the CPU path on a strided target subset against ALL sources (onbody's ntskip
    method, reference extern/onbody/src/main2dvort.cpp:116-133), sized for ~`seconds`."""
    import oracle
    panels = W.multi_circle_panels(*wl["panels"]) if wl.get("panels") else None
    kind, fn = cpu_checker(wl["core"], panels)
    oracle.set_threads()          # torchrun exports OMP_NUM_THREADS=1; use every host core
    sx, sy, ss, sr = parts
    n = sx.size
    # velocity + gradient: the reference defines the pair kernel (kernelug_0v_0b) but no loop drives it;
    # the CPU arm is that kernel inside the reference's own loop shape, OpenMP over targets (oracle/ref_driver.cpp)
    res = wl["results"]

    def run(count):
        sub = W.strided_subset(n, count)
        t0 = time.perf_counter()
        out = fn(res, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub] if wl["blob"] else None)
        return time.perf_counter() - t0, sub, out

    cal = min(n, max(64, int(2e8 // max(n, 1))))
    run(cal)                                  # warm-up (thread team start-up, page faults)
    cal = min(n, 8 * cal)
    dt, _, _ = run(cal)                       # calibration
    rate = cal * n / max(dt, 1e-6)
    count = int(min(n, max(cal, seconds * rate / n)))
    return kind, run, count


def sources_per_target(wl, n):
    return n + (wl["panels"][0] * wl["panels"][1] if wl.get("panels") else 0)


def run_reference_arm(args, wl, n):
    """--impl reference: the reference's OpenMP direct sum on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    parts = make_particles(args.workload, n)
    kind, run, count = cpu_sample(wl, parts, args.cpu_seconds)
    for _ in range(max(0, args.warmup - 1)):   # calibration already ran the path once
        run(min(count, 256))
    times = []
    for _ in range(args.steps):
        dt, _, _ = run(count)
        times.append(dt)
    total = sum(times)
    value = args.steps * count * sources_per_target(wl, n) / total
    import oracle
    cores = oracle.host_cores()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 pair terms, f64 accumulation",
        "data": "synthetic",
        # same keys as our arm's config (the driver compares the two)
        "config": {"workload": wl["desc"], "particles": n, "particles_per_gpu": sharding.block_size(n, max(1, int(os.environ.get("WORLD_SIZE", "1")))),
                   "core": wl["core"], "results": wl["results"],
                   "pairs_per_step": float(n) * float(n) + (float(n) * wl["panels"][0] * wl["panels"][1] if wl.get("panels") else 0.0),
                   "l2": "n/a (CPU arm)", "parallelism": "OpenMP over targets on the host cores (rank 0 only)",
                   "sample": f"{count} strided targets x {n} sources per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{count} strided targets x all {n} sources per step, OpenMP on all host cores, USE_VC off"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.path = tempfile.mktemp(prefix="o2d_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(device_index)],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # keep samples taken under load (top half of power draw)
            thr = np.median(pw)
            load = [s for s, w in zip(sm, pw) if w >= thr]
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
CPU_FLAGS_NOTE = ("reference headers compiled by oracle/build_ref.sh: g++ -O3 -fopenmp -march=x86-64-v3 (portable between "
                  "boxes; the reference's own flag is -march=native), USE_VC off (Vc is not vendored)")


def gather_subset(tensors, lo, hi, sub, world):
    """Values of block-sharded device tensors at the global indices `sub` (sorted), assembled on every rank."""
    import torch.distributed as dist
    import torch
    mine = sub[(sub >= lo) & (sub < hi)] - lo
    idx = torch.from_numpy(mine).to(tensors[0].device)
    vals = [t[idx].cpu().numpy() for t in tensors]
    if world == 1:
        return vals
    parts = [None] * world
    dist.all_gather_object(parts, vals)
    return [np.concatenate([p[k] for p in parts]) for k in range(len(tensors))]


def panel_parity(ctx, panels, tx, ty):
    """The panel-kernel tolerance question in the open (README "Parity rule for panels"): the CUDA panels-on-points result
    on the given targets against the reference's float32 loop AND against float64, with the reference's own distance
    from float64 beside them."""
    import oracle
    from oracle.f64 import panels_on_points_f64
    nx, ny, idx, vs, ss = panels
    got = ctx.panels_on_points_np(nx, ny, idx, vs, ss, tx, ty)
    chk = oracle.ref("rm") if oracle.have_ref() else oracle.port()
    want = chk.panels_on_points(nx, ny, idx, vs, ss, tx, ty)
    exact = panels_on_points_f64(nx, ny, idx, vs, ss, tx, ty)
    ref_noise = W.rel_err(want, exact)
    e_ref, e_f64 = W.rel_err(got, want), W.rel_err(got, exact)
    ok = e_ref <= 1e-5 or (e_f64 <= 1e-5 and e_ref <= 1.5 * ref_noise)
    return {"targets": int(tx.size), "panels": int(vs.size), "err_vs_reference": e_ref, "err_vs_f64": e_f64,
            "reference_err_vs_f64": ref_noise, "within_rule": bool(ok),
            "rule": "err_vs_reference <= 1e-5, or (reference_err_vs_f64 > 1e-5 and err_vs_f64 <= 1e-5 and "
                    "err_vs_reference <= 1.5 x reference_err_vs_f64)"}


def strong_block(args, ctx, dev, stream, flush, world, rank, barrier):
    """BASELINE configs[3] at FIXED size next to the weak-scaling headline: one find_vels (zero, particles-on-particles,
    panels-on-particles, finalize) of 4M particles + 20k panels through o2d_find_vels_sharded, so that the per-N lines of a
    scaling run carry a strong-scaling curve with its own parity number."""
    import torch
    import torch.distributed as dist
    import omega2d_b200 as o2d
    wl = WORKLOADS["c4"]
    n = args.strong_particles or wl["n"]
    lo, hi = sharding.block_range(n, world, rank)
    nblock = sharding.block_size(n, world)
    nloc = hi - lo
    parts = make_particles("c4", n)
    panels = W.multi_circle_panels(*wl["panels"])
    d = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in parts]
    pd = [torch.from_numpy(a).to(dev) for a in (panels[0], panels[1], panels[2].astype(np.int32), panels[3], panels[4])]
    npan, nn = panels[3].size, panels[0].size
    tu = torch.zeros(max(nloc, 1), device=dev); tv = torch.zeros(max(nloc, 1), device=dev)

    def step():
        ctx.find_vels_sharded(wl["core"], wl["results"], o2d.MEM_DEVICE, nloc, nblock, d[0], d[1], d[2], d[3],
                              npan, nn, pd[0], pd[1], pd[2], pd[3], pd[4], nloc, d[0], d[1], d[3], tu, tv, None, (0.0, 0.0))
    step()
    barrier()
    steps = max(1, args.strong_steps)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(float(k))
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    pairs = float(n) * float(n) + float(n) * npan
    sub = W.strided_subset(n, 1024)
    gu, gv = gather_subset([tu, tv], lo, hi, sub, world)
    if rank != 0:
        return None
    out = {"workload": wl["desc"], "scaling": "strong", "particles": n, "panels": int(npan), "particles_per_gpu": nblock,
           "steps": steps, "warmup": 1, "ms_per_step": ms, "value": pairs / (ms * 1e-3), "unit": UNIT,
           "pairs_per_step": pairs, "api": "o2d_find_vels_sharded(O2D_MEM_DEVICE)"}
    if not args.no_cpu_baseline:
        import oracle
        oracle.set_threads()
        kind, fn = cpu_checker(wl["core"], panels)
        sx, sy, ss, sr = parts
        t0 = time.perf_counter()
        ru, rv = fn(wl["results"], sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])[:2]
        dt = time.perf_counter() - t0
        out["max_rel_err_gpu_vs_cpu"] = W.rel_err((gu, gv), (ru, rv))
        out["tolerance"] = 1e-5
        # the panel term of the reference is float32 atan2f-atan2f / logf (src/Kernels.h:262-268,338) and is itself
        # further than 1e-5 from exact arithmetic for these short panels (panels_only below), so the total is also
        # compared with the reference's own total after its panel term is replaced by the float64 evaluation
        from oracle.f64 import panels_on_points_f64
        chk = oracle.ref(wl["core"]) if oracle.have_ref() else oracle.port()
        rp = chk.panels_on_points(*panels, sx[sub], sy[sub])
        ep = panels_on_points_f64(*panels, sx[sub], sy[sub])
        k = 0.5 / np.pi
        fixed = (ru.astype(np.float64) + k * (ep[0] - rp[0]), rv.astype(np.float64) + k * (ep[1] - rp[1]))
        out["max_rel_err_vs_reference_with_float64_panel_term"] = W.rel_err((gu, gv), fixed)
        out["reference_total_err_vs_its_float64_panel_form"] = W.rel_err((ru, rv), fixed)
        out["cpu_baseline"] = {"value": sub.size * (float(n) + npan) / dt, "unit": UNIT, "cores": oracle.host_cores(), "kind": kind,
                               "sample": f"find_vels on {sub.size} strided targets x all {n} particles + {npan} panels, one call; " + CPU_FLAGS_NOTE}
        out["panels_only"] = panel_parity(ctx, panels, sx[sub], sy[sub])
    return out


def run_ours(args, wl, n1):
    import torch
    import torch.distributed as dist
    import omega2d_b200 as o2d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; omega2d_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE JSON line: whatever libraries print while communicators come up (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr at the file-descriptor level until the line is due
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def restore_stdout():
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = n1 if (world == 1 or args.scaling == "strong") else sharding.weak_scaling_count(n1, world)
    lo, hi = sharding.block_range(n, world, rank)
    nblock = sharding.block_size(n, world)
    nloc = hi - lo
    parts = make_particles(args.workload, n)          # every rank generates the same seeded set
    sx, sy, ss, sr = parts
    core, results = wl["core"], wl["results"]
    nout = {1: 2, 2: 6, 4: 3, 5: 3}[results]

    ctx = o2d.Context(local_rank)
    if world > 1:
        uid = [o2d.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    # a dedicated (non-default) stream shared by torch and the library: torch's default stream has
    # handle 0, which o2d_set_stream reads as "use the context's own stream"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    info = ctx.device_info()

    # ---- device-resident inputs (this rank's block) --------------------------------------------
    d = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (sx, sy, ss, sr)]
    outs = [torch.zeros(max(nloc, 1), device=dev) for _ in range(nout)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2

    pan = None
    if wl.get("panels"):
        pnx, pny, pidx, pvs, pss = W.multi_circle_panels(*wl["panels"])
        pan = dict(np=pvs.size, nn=pnx.size, host=(pnx, pny, pidx, pvs, pss),
                   dev=[torch.from_numpy(a).to(dev) for a in (pnx, pny, pidx.astype(np.int32), pvs, pss)])

    def step_device():
        tq = outs[2] if nout == 3 else None
        g = outs[2:6] if nout == 6 else [None] * 4
        tr = d[3] if wl["blob"] else None
        if pan is not None:
            p = pan["dev"]
            ctx.find_vels_sharded(core, results, o2d.MEM_DEVICE, nloc, nblock, d[0], d[1], d[2], d[3],
                                  pan["np"], pan["nn"], p[0], p[1], p[2], p[3], p[4],
                                  nloc, d[0], d[1], tr, outs[0], outs[1], tq, (0.0, 0.0))
        elif world == 1:
            ctx.points_on_points(core, results, o2d.MEM_DEVICE, nloc, d[0], d[1], d[2], d[3],
                                 nloc, d[0], d[1], tr, outs[0], outs[1], tq, *g)
        else:
            ctx.points_on_points_sharded(core, results, o2d.MEM_DEVICE, nloc, nblock, d[0], d[1], d[2], d[3],
                                         nloc, d[0], d[1], tr, outs[0], outs[1], tq, *g)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        for o in outs:
            o.zero_()
        step_device()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream -----------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.reset_counters()
    ctx.enable_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(float(k))                 # evict L2 between timed iterations (outside the event pair)
        for o in outs:
            o.zero_()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    if os.environ.get("O2D_BENCH_DEBUG"):
        print(f"[rank {rank}] per-step ms: " + " ".join(f"{a.elapsed_time(b):.2f}" for a, b in ev), file=sys.stderr, flush=True)
    cnt = ctx.counters()
    ctx.enable_timing(False)
    t = torch.tensor([ms, cnt["pair_ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, pair_ms_max = float(t[0]), float(t[1])
    pairs_per_step = float(n) * float(n) + (float(n) * pan["np"] if pan is not None else 0.0)
    value = args.steps * pairs_per_step / (ms_max * 1e-3)

    # ---- results of this rank's block (checked further down; oracle as checker only) -------------------
    u = outs[0][:nloc].cpu().numpy(); v = outs[1][:nloc].cpu().numpy()

    # ---- e2e: HOST pinned buffers through the C ABI, H2D + D2H inside the timed region ------------------
    e2e = None
    if not args.no_e2e:
        h = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).pin_memory() for a in (sx, sy, ss, sr)]
        houts = [torch.zeros(max(nloc, 1)).pin_memory() for _ in range(nout)]

        def step_host():
            tq = houts[2] if nout == 3 else None
            g = houts[2:6] if nout == 6 else [None] * 4
            tr = h[3] if wl["blob"] else None
            if pan is not None:
                p = pan["host"]
                ctx.find_vels_sharded(core, results, o2d.MEM_HOST, nloc, nblock, h[0], h[1], h[2], h[3],
                                      pan["np"], pan["nn"], p[0], p[1], p[2], p[3], p[4],
                                      nloc, h[0], h[1], tr, houts[0], houts[1], tq, (0.0, 0.0))
            elif world == 1:
                ctx.points_on_points(core, results, o2d.MEM_HOST, nloc, h[0], h[1], h[2], h[3],
                                     nloc, h[0], h[1], tr, houts[0], houts[1], tq, *g)
            else:
                ctx.points_on_points_sharded(core, results, o2d.MEM_HOST, nloc, nblock, h[0], h[1], h[2], h[3],
                                             nloc, h[0], h[1], tr, houts[0], houts[1], tq, *g)

        for o in houts:
            o.zero_()
        step_host()                                            # warm the staging buffers
        barrier()
        e_steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            for o in houts:
                o.zero_()
            step_host()                                        # synchronous: returns after the D2H
        torch.cuda.synchronize()
        et = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        # the same call on PAGEABLE arrays -- what the reference's std::vector storage is
        ph = [np.ascontiguousarray(a[lo:hi]).copy() for a in (sx, sy, ss, sr)]
        pouts = [np.zeros(max(nloc, 1), np.float32) for _ in range(nout)]

        def step_pageable():
            tq = pouts[2] if nout == 3 else None
            g = pouts[2:6] if nout == 6 else [None] * 4
            tr = ph[3] if wl["blob"] else None
            if pan is not None:
                p = pan["host"]
                ctx.find_vels_sharded(core, results, o2d.MEM_HOST, nloc, nblock, ph[0], ph[1], ph[2], ph[3],
                                      pan["np"], pan["nn"], p[0], p[1], p[2], p[3], p[4],
                                      nloc, ph[0], ph[1], tr, pouts[0], pouts[1], tq, (0.0, 0.0))
            elif world == 1:
                ctx.points_on_points(core, results, o2d.MEM_HOST, nloc, ph[0], ph[1], ph[2], ph[3],
                                     nloc, ph[0], ph[1], tr, pouts[0], pouts[1], tq, *g)
            else:
                ctx.points_on_points_sharded(core, results, o2d.MEM_HOST, nloc, nblock, ph[0], ph[1], ph[2], ph[3],
                                             nloc, ph[0], ph[1], tr, pouts[0], pouts[1], tq, *g)
        step_pageable()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            for o in pouts:
                o[:] = 0.0
            step_pageable()
        torch.cuda.synchronize()
        pt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        h2d = 4 * (4 * nloc + (3 if wl["blob"] else 2) * nloc + nout * nloc) * world
        if pan is not None:   # find_vels overwrites its outputs (no upload of them); panels go up on every rank
            h2d = (4 * (4 * nloc + (3 if wl["blob"] else 2) * nloc) + 4 * (2 * pan["nn"] + 4 * pan["np"])) * world
        d2h = 4 * nout * nloc * world
        e2e = {"value": e_steps * pairs_per_step / float(et[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e_steps,
               "value_pageable_arrays": e_steps * pairs_per_step / float(pt[0]),
               "api": ("o2d_find_vels_sharded" if pan is not None else "o2d_points_on_points[_sharded]") +
                      "(O2D_MEM_HOST) on pinned host arrays"}
        # host path and device path must agree bit for bit
        if not np.array_equal(houts[0][:nloc].numpy(), u):
            bad = int((houts[0][:nloc].numpy() != u).sum())
            raise SystemExit(f"bench.py: host-memory path disagrees with device path on {bad} of {nloc} targets")

    # ---- parity sample: a strided subset of ALL targets, gathered from the ranks that own them ---------------------
    cpu_plan = None
    gsub = None
    if not args.no_cpu_baseline:
        plan = [None]
        if rank == 0:
            kind, run, count = cpu_sample(wl, parts, args.cpu_seconds if world == 1 else min(args.cpu_seconds, 6.0))
            plan = [(kind, count)]
            cpu_plan = (kind, run, count)
        if world > 1:
            dist.broadcast_object_list(plan, src=0)
        sub_idx = W.strided_subset(n, plan[0][1])
        gsub = gather_subset(outs[:nout], lo, hi, sub_idx, world)

    # ---- fixed-size C4 next to the weak-scaling headline (default workload only) -------------------------------------
    strong = None
    if args.workload == "c3" and not args.no_strong:
        d.clear()
        strong = strong_block(args, ctx, dev, stream, flush, world, rank, barrier)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        restore_stdout()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    flops_pp = o2d.flops_per_pair(0, core, results, wl["blob"])
    pairs_rank0 = float(nloc) * float(n)
    panel_flops = 47.0 * float(nloc) * pan["np"] if pan is not None else 0.0   # flopsu_1vs_0p (src/Kernels.h:362)
    kern_s = pair_ms_max * 1e-3 / args.steps          # pair-sum kernel time per step (max over ranks)
    achieved = (flops_pp * pairs_rank0 + panel_flops) / kern_s / 1e12 if kern_s > 0 else 0.0
    peak_nominal = info["fp32_peak_tflops"]
    ffma = ctx.measure_fp32_peak(False)
    ffma2 = ctx.measure_fp32_peak(True)
    traffic = ncu_traffic(core, results, n) if world == 1 else None
    # hardware view: FP32 lane-ops per pair COUNTED by ncu on this kernel (FMA-pipe warp instructions x 64 lanes, capture at
    # 1M x 1M in profiles/) x the pairs of this launch / the live kernel time, against what the FMA pipe retires
    # (128 lane-ops per clock per SM); the XU (MUFU) pipe likewise against 16 per clock per SM
    cap = ncu_traffic(core, results, 1_000_000)
    hw = None
    if cap and cap.get("fp32_lane_ops_per_launch") and kern_s > 0:
        clk = info["sm_clock_mhz"] * 1e6 * info["sm_count"]
        lane_pp, mufu_pp = cap["fp32_lane_ops_per_launch"] / 1e12, cap["mufu_lane_ops_per_launch"] / 1e12
        hw = {"frac_hw": lane_pp * pairs_rank0 / kern_s / (128.0 * clk), "xu_frac_hw": mufu_pp * pairs_rank0 / kern_s / (16.0 * clk),
              "fp32_lane_ops_per_pair": lane_pp, "mufu_per_pair": mufu_pp,
              "fma_pipe_cycles_active_pct_under_ncu": cap["fma_pipe_cycles_active_pct"], "xu_pipe_pct_under_ncu": cap["xu_pipe_pct"],
              "issue_active_pct_under_ncu": cap["issue_active_pct"], "source": cap["source"]}
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
        "frac": achieved / peak_nominal if peak_nominal else None,
        "frac_hw": hw["frac_hw"] if hw else None, "hardware": hw,
        "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
        "algorithmic_bytes_per_launch": 16 * n + 12 * nloc + 4 * nout * nloc,
        "flops_per_pair": flops_pp, "kernel": f"pp_kernel<{core},res{results}>",
        "kernel_ms_per_step": kern_s * 1e3,
        "peak_source": f"nominal: {info['sm_count']} SMs x 128 lanes x 2 x {info['sm_clock_mhz']} MHz "
                       "(MEASURED_PEAKS.json has no FP32 non-tensor entry; its hbm/bf16 peaks do not bound this kernel)",
        "measured_ffma_tflops": ffma, "measured_ffma2_tflops": ffma2,
        "frac_of_measured_ffma2": achieved / ffma2 if ffma2 else None,
        "note": "`frac` counts the reference's own per-pair flops (src/Kernels.h, src/CoreFunc.h: 34 for velocity+gradient, where the "
                "kernel executes 15 FP32 lane-ops + 1 MUFU), so it is an algorithmic rate, not hardware utilisation; `frac_hw` is the "
                "FMA-pipe utilisation from ncu-counted instructions",
    }

    # ---- other kernels of the path on the same particle count (device-resident, kernel-only view) -----------
    variants = None
    if world == 1 and not args.no_variants and args.workload == "c3":
        variants = []
        for vname, vcore, vres, vary in (("velocity only (kernelu_0v_0b: what find_vels evaluates every substep), per-particle radii", "rm", 1, True),
                                         ("velocity only, uniform radius 1.5/sqrt(N) (the reference's normal case)", "rm", 1, False),
                                         ("velocity only, Vatistas-2 core (reference default), uniform radius", "v2", 1, False)):
            vp = make_particles(args.workload, n, vary_radius=vary)
            vd = [torch.from_numpy(a).to(dev) for a in vp]
            vn = {1: 2, 2: 6}[vres]
            vo = [torch.zeros(n, device=dev) for _ in range(vn)]

            def vstep():
                for o in vo:
                    o.zero_()
                ctx.points_on_points(vcore, vres, o2d.MEM_DEVICE, n, *vd, n, vd[0], vd[1], vd[3], vo[0], vo[1], None,
                                     *(vo[2:6] if vn == 6 else [None] * 4))
            vstep()
            torch.cuda.synchronize()
            flush.fill_(1.0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            vsteps = 2
            e0.record(stream)
            for _ in range(vsteps):
                vstep()
            e1.record(stream)
            torch.cuda.synchronize()
            vms = e0.elapsed_time(e1) / vsteps
            fpp = o2d.flops_per_pair(0, vcore, vres, True)
            rate = float(n) * float(n) / (vms * 1e-3)
            variants.append({"what": vname, "core": vcore, "results": vres, "ms_per_step": vms, "value": rate,
                             "flops_per_pair": fpp, "frac_fp32_peak": fpp * rate / 1e12 / peak_nominal})
            del vd, vo

    # ---- C1, the reference's own CPU-runnable case, as a latency config: one whole Ralston RK2 convection step
    #      (2 evaluations of 20k x 20k pairs) through o2d_advect with HOST arrays, next to the CPU reference ----
    c1 = None
    if world == 1 and not args.no_variants and args.workload == "c3":
        cx, cy, cs, cr = W.c1_speedtest(20000)
        # pageable host arrays, like the reference's std::vector storage
        hx = [a.copy() for a in (cx, cy, cs, cr)]
        hu = [np.zeros(20000, np.float32) for _ in range(2)]

        def c1_run(nsteps):
            hx[0][:] = cx; hx[1][:] = cy
            ctx.advect("v2", 2, o2d.MEM_HOST, 20000, hx[0], hx[1], hx[2], hx[3], hu[0], hu[1], dt=0.005, nsteps=nsteps)
        for _ in range(3):
            c1_run(1)
        t0 = time.perf_counter()
        for _ in range(10):
            c1_run(1)
        gpu_ms = (time.perf_counter() - t0) / 10 * 1e3
        c1_run(10)
        t0 = time.perf_counter()
        for _ in range(3):
            c1_run(10)
        run_ms = (time.perf_counter() - t0) / 3 * 1e3
        c1 = {"what": "C1 speedtest20kx10 restated (20k blobs, V2 core, inviscid, Ralston RK2, dt 0.005; 2 evaluations of 4e8 pairs "
                      "per step), pageable host arrays in and out: gpu_ms_per_step = one o2d_advect call per step (upload, both "
                      "stages on the device, download); gpu_ms_10_steps = the whole 10-step run as ONE o2d_advect_steps call "
                      "(particles resident on the GPU between steps)",
              "gpu_ms_per_step": gpu_ms, "gpu_ms_10_steps": run_ms, "pairs_per_step": 2 * 20000.0 ** 2,
              "gpu_value": 2 * 20000.0 ** 2 / (gpu_ms * 1e-3), "gpu_value_10_steps": 20 * 20000.0 ** 2 / (run_ms * 1e-3), "unit": UNIT}
        if not args.no_cpu_baseline:
            import oracle
            oracle.set_threads()
            step = (lambda x, y: oracle.ref("v2").advect(2, x, y, cs, cr, 0.005)) if oracle.have_ref() else \
                   (lambda x, y: oracle.port().advect("v2", 2, x, y, cs, cr, 0.005))
            step(cx, cy)
            t0 = time.perf_counter()
            wx, wy = cx, cy
            for _ in range(10):
                wx, wy, _, _ = step(wx, wy)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            derr = float(np.abs(np.stack([hx[0] - wx, hx[1] - wy])).max())
            c1.update(cpu_ms_10_steps=cpu_ms, cpu_ms_per_step=cpu_ms / 10, cpu_value=20 * 20000.0 ** 2 / (cpu_ms * 1e-3),
                      cpu_cores=oracle.host_cores(), cpu_kind="reference" if oracle.have_ref() else "port",
                      max_abs_position_diff_after_10_steps=derr)

    # ---- parity spot check + CPU baseline (oracle used as checker / reported baseline only), at every N -----------------
    cpu = None
    if cpu_plan is not None:
        kind, run, count = cpu_plan
        dt, sub, ref_out = run(count)
        gu, gv = gsub[0].astype(np.float64), gsub[1].astype(np.float64)
        # max|u| over the sample (a strided sample of a Gaussian blob sees the same maximum as the full set to <1 %)
        scale = np.sqrt(np.asarray(ref_out[0], np.float64) ** 2 + np.asarray(ref_out[1], np.float64) ** 2).max()
        err = np.sqrt((gu - ref_out[0]) ** 2 + (gv - ref_out[1]) ** 2).max() / scale
        if results == 2:      # every gradient component, same form of the tolerance
            for k in range(2, 6):
                err = max(err, float(np.abs(gsub[k].astype(np.float64) - ref_out[k]).max() / np.abs(np.asarray(ref_out[k], np.float64)).max()))
        # the two numbers the reference's nearest thing to a correctness test prints (onbody, extern/onbody/src/
        # main2dvort.cpp:139-154; u component only): sqrt(max du^2 / mean u^2) and sqrt(sum du^2 / sum u^2)
        du2 = (gu - np.asarray(ref_out[0], np.float64)) ** 2
        u2 = np.asarray(ref_out[0], np.float64) ** 2
        onbody = {"max_error": float(np.sqrt(du2.max() / u2.mean())), "rms_error": float(np.sqrt(du2.sum() / u2.sum()))}
        import oracle
        cpu = {"value": count * sources_per_target(wl, n) / dt, "unit": UNIT, "cores": oracle.host_cores(), "kind": kind,
               "sample": f"{count} strided targets x all {n} sources, one call, OpenMP on all host cores; " + CPU_FLAGS_NOTE
                         + ("; the reference's kernelug_0v_0b in its own loop shape (it ships no velandgrad loop)" if results == 2 else ""),
               "seconds": dt, "max_rel_err_gpu_vs_cpu": float(err), "tolerance": 1e-5, "onbody_error_metrics_u": onbody}
        if pan is not None:
            cpu["panels_only"] = panel_parity(ctx, pan["host"], sx[sub], sy[sub])

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32 pair terms, f64 accumulation", "data": "synthetic",
        "config": {"workload": wl["desc"], "particles": n, "particles_per_gpu": nblock, "core": core,
                   "results": results, "pairs_per_step": pairs_per_step, "l2": "flushed between timed iterations (256 MB write)",
                   "sample": "every target x every source, every step",
                   "parallelism": "single GPU" if world == 1 else f"targets sharded in {world} contiguous blocks, "
                                  "sources all-gathered (NCCL) once per step"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": cnt["launches"],
        "variants": variants, "c1_step": c1, "strong": strong,
        "clocks": clocks, "wall_s_timed_region": wall,
        "frac_fp32_peak_whole_step": (flops_pp * float(n) * float(n) + panel_flops * world) / (ms_max * 1e-3 / args.steps)
                                     / world / 1e12 / peak_nominal,
    }
    restore_stdout()
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()
        restore_stdout()


def main():
    args = parse_args()
    wl = WORKLOADS[args.workload]
    if args.scaling is None:
        args.scaling = wl.get("scaling", "weak")
    n1 = args.n or wl["n"]
    if args.impl == "reference":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        n = n1 if (world == 1 or args.scaling == "strong") else sharding.weak_scaling_count(n1, world)
        run_reference_arm(args, wl, n)
        return
    run_ours(args, wl, n1)


if __name__ == "__main__":
    main()
