"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- contiguous block
partition, per-rank block packing to a common capacity, all-gather, and the equivalence
'sum over gathered sources on each rank's targets == single-process sum' -- exercised with the
oracle standing in for the device kernel (this tests the sharding logic, not CUDA)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omega2d_b200 import sharding, workloads as W


def test_block_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 1_000_003):
        for p in (1, 2, 3, 8):
            b = sharding.block_size(n, p)
            spans = [sharding.block_range(n, p, r) for r in range(p)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            assert all(0 <= hi - lo <= b for lo, hi in spans)
    assert sharding.weak_scaling_count(1_000_000, 4) == 2_000_000
    assert sharding.weak_scaling_count(1_000_000, 8) % 8 == 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port_no, n, ret):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sx, sy, ss, sr = W.onbody_random(n, seed=77)
    lo, hi = sharding.block_range(n, world, rank)
    nb = sharding.block_size(n, world)
    # the layout of o2d_pp_sharded_device: every rank packs its block into its slot of `blk` entries (padding as
    # pack_sources_kernel writes it), one all-gather replicates the slots, and each rank pads the tail of the last
    # tile itself -- then ONE sum over the whole gathered array
    blk, total = sharding.gather_layout(nb, world)
    assert blk % 2 == 0 and blk >= nb and total % 1024 == 0 and total >= blk * world
    k = hi - lo
    local = [a[lo:hi] for a in (sx, sy, ss, sr)]
    mine = np.stack(sharding.pad_block(local, blk))
    gathered = [torch.zeros(4, blk) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    tail = np.stack(sharding.pad_block(local, k + (total - blk * world)))[:, k:]
    allsrc = np.concatenate([torch.cat(gathered, dim=1).numpy(), tail], axis=1)
    assert allsrc.shape[1] == total
    P = oracle.port()
    got = P.points_on_points("v2", oracle.VELONLY, allsrc[0], allsrc[1], allsrc[2], allsrc[3], sx[lo:hi], sy[lo:hi], sr[lo:hi])
    want = P.points_on_points("v2", oracle.VELONLY, sx, sy, ss, sr, sx[lo:hi], sy[lo:hi], sr[lo:hi])
    err = W.rel_err(got, want) if k else 0.0
    t = torch.tensor([err], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret.put(float(t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 1])
def test_two_rank_gather_equals_single_process(n):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # padding adds exact zeros, so the sums agree to double-accumulation order (bitwise in practice)
    assert ret.get(timeout=5) < 1e-7


def _panel_worker(rank, world, port_no, n, ret):
    """BEM right-hand side with sharded particles: per-rank reciprocity sums, all-reduce, same result everywhere
    (the host-side logic of o2d_points_on_panels_sharded)."""
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, s, _ = W.annulus_blobs(n, seed=12)
    nx, ny, idx, _, _ = W.circle_panels(64, radius=0.5)
    lo, hi = sharding.block_range(n, world, rank)
    P = oracle.port()
    area = P.panel_areas(nx, ny, idx)
    if hi > lo:
        pu, pv = P.points_on_panels(x[lo:hi], y[lo:hi], s[lo:hi], nx, ny, idx, area)   # from zero: -(local sum)/area
    else:
        pu, pv = np.zeros(64, np.float32), np.zeros(64, np.float32)
    t = torch.from_numpy(np.stack([pu, pv]).astype(np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    want = P.points_on_panels(x, y, s, nx, ny, idx, area)
    err = W.rel_err((t[0].numpy(), t[1].numpy()), want)
    e = torch.tensor([err], dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret.put(float(e))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [777, 1])
def test_two_rank_panel_sums_allreduce(n):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_panel_worker, args=(r, 2, port_no, n, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) < 1e-6
