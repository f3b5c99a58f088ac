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
t = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(t) else "FAIL", flush=True)
dist.barrier(); dist.destroy_process_group()
