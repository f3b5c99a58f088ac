"""bench.py prints exactly one JSON line with the keys the driver reads -- checked on a reduced particle count
(--n) so that it runs in seconds: the reference arm on the CPU (not gpu), our arm on the GPU (gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # ONE line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--n", "6000", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.3")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "biot_savart_pair_interactions_per_sec" and d["unit"] == "interactions/s" and d["higher_is_better"] is True
    assert d["value"] > 1e6 and d["steps"] == 2 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--n", "40000", "--steps", "3", "--warmup", "3", "--cpu-seconds", "0.5", "--no-variants", "--strong-particles", "50000")
    assert BASE | {"roofline", "clocks"} <= set(d) and "impl" not in d
    assert d["metric"] == "biot_savart_pair_interactions_per_sec" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["value"] > 1e11 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0.0 < r["frac"] < 1.5 and r["unit"] == "TFLOP/s"
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["max_rel_err_gpu_vs_cpu"] < 1e-5
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["config"]["results"] == 2 and "velocity + gradient" in d["config"]["workload"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert e["value_pageable_arrays"] > 0
    # the fixed-size C4 block beside the headline: its own value, parity vs the reference's find_vels, and the panel rule
    s = d["strong"]
    assert s["scaling"] == "strong" and s["value"] > 0 and s["particles"] == 50000 and s["panels"] == 20000
    assert s["max_rel_err_vs_reference_with_float64_panel_term"] < 1e-5 and s["panels_only"]["within_rule"] is True
    assert s["panels_only"]["err_vs_f64"] < 1e-5


def _gather_worker(rank, world, port_no, ret):
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    from omega2d_b200 import sharding, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1001
    lo, hi = sharding.block_range(n, world, rank)
    a = torch.arange(lo, hi, dtype=torch.float32)
    sub = W.strided_subset(n, 37)
    got = bench.gather_subset([a, 2 * a], lo, hi, sub, world)
    ok = np.array_equal(got[0], sub.astype(np.float32)) and np.array_equal(got[1], 2 * sub.astype(np.float32))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret.put(float(t))
    dist.barrier()
    dist.destroy_process_group()


def test_parity_sample_is_gathered_from_the_ranks_that_own_it():
    """world_size 2 over gloo: the strided parity sample of an N>1 bench line comes back in global order on every rank."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port_no, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == 1.0
