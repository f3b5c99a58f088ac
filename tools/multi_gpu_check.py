"""torchrun --nproc-per-node P tools/multi_gpu_check.py  -- sharded path vs oracle (rank-local strided subset)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import omega2d_b200 as o2d, oracle
from omega2d_b200 import sharding, workloads as W

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = o2d.Context(lr)
uid = [o2d.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(world, rank, uid[0])
ok = True
for n, core, res in ((100_003, "v2", o2d.VELONLY), (250_000, "rm", o2d.VELANDGRAD), (4099, "wl", o2d.VELANDVORT), (world - 1 if world > 1 else 1, "rm", o2d.VELONLY)):
    sx, sy, ss, sr = W.gaussian_blob(max(n, 1), seed=n)
    lo, hi = sharding.block_range(n, world, rank)
    nb = sharding.block_size(n, world)
    nloc = hi - lo
    nout = {1: 2, 2: 6, 4: 3}[res]
    outs = [np.zeros(nloc, np.float32) for _ in range(nout)]
    tq = outs[2] if nout == 3 else None
    g = outs[2:6] if nout == 6 else [None] * 4
    ctx.points_on_points_sharded(core, res, o2d.MEM_HOST, nloc, nb, sx[lo:hi], sy[lo:hi], ss[lo:hi], sr[lo:hi],
                                 nloc, sx[lo:hi], sy[lo:hi], sr[lo:hi], outs[0], outs[1], tq, *g)
    if nloc > 0:
        sub = lo + W.strided_subset(nloc, 64)
        P = oracle.port()
        if res == o2d.VELANDGRAD:
            want = P.points_on_points_grad(core, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])
        else:
            want = P.points_on_points(core, res, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])
        errs = []
        for k in range(nout):
            sc = max(np.abs(want[k]).max(), 1e-30)
            errs.append(float(np.abs(outs[k][sub - lo] - want[k]).max() / sc))
        good = max(errs) < 2e-5
        ok &= good
        print(f"rank {rank} n={n} {core} res={res} nloc={nloc} max err {max(errs):.2e} {'OK' if good else 'FAIL'}", flush=True)
    else:
        print(f"rank {rank} n={n}: empty block OK", flush=True)
# sharded find_vels (particles sharded, panels replicated) and sharded points-on-panels (all-reduce)
n = 60_001
sx, sy, ss, sr = W.gaussian_blob(n, seed=77, signed=True)
ss = (ss * n).astype(np.float32)
nx, ny, idx, pvs, pss = W.circle_panels(257, radius=0.3, center=(0.9, 0.2))
lo, hi = sharding.block_range(n, world, rank); nb = sharding.block_size(n, world); nloc = hi - lo
tu = np.empty(nloc, np.float32); tv = np.empty(nloc, np.float32); tw = np.empty(nloc, np.float32)
ctx.find_vels_sharded("v2", o2d.VELANDVORT, o2d.MEM_HOST, nloc, nb, sx[lo:hi], sy[lo:hi], ss[lo:hi], sr[lo:hi],
                      pvs.size, nx.size, nx, ny, idx, pvs, pss, nloc, sx[lo:hi], sy[lo:hi], sr[lo:hi], tu, tv, tw, (1.0, -0.5))
P = oracle.port()
if nloc > 0:
    sub = lo + W.strided_subset(nloc, 48)
    want = P.find_vels("v2", oracle.VELANDVORT, (sx, sy, ss, sr), (nx, ny, idx, pvs, pss), sx[sub], sy[sub], sr[sub], (1.0, -0.5))
    e = max(W.rel_err((tu[sub - lo], tv[sub - lo]), want[:2]), W.rel_err(tw[sub - lo], want[2]))
    good = e < 1e-5
    ok &= good
    print(f"rank {rank} find_vels_sharded max err {e:.2e} {'OK' if good else 'FAIL'}", flush=True)
area = P.panel_areas(nx, ny, idx)
pu = np.full(pvs.size, 0.25, np.float32); pv = np.full(pvs.size, -0.5, np.float32)
ctx.points_on_panels_sharded(o2d.MEM_HOST, nloc, sx[lo:hi], sy[lo:hi], ss[lo:hi], pvs.size, nx.size, nx, ny, idx, area, pu, pv)
want = P.points_on_panels(sx, sy, ss, nx, ny, idx, area, np.full(pvs.size, 0.25, np.float32), np.full(pvs.size, -0.5, np.float32))
e = W.rel_err((pu, pv), want)
good = e < 1e-5
ok &= good
print(f"rank {rank} points_on_panels_sharded max err {e:.2e} {'OK' if good else 'FAIL'}", flush=True)

t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(t) else "FAIL", flush=True)
dist.barrier(); dist.destroy_process_group()
