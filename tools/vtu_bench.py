"""One o2d_write_vtu_points call on 1M device-resident particles (ncu target: vtu_encode_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import omega2d_b200 as o2d

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = o2d.Context(0)
rng = np.random.default_rng(0)
d = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)).cuda() for _ in range(6)]
for _ in range(3):
    ctx.write_vtu_points(o2d.MEM_DEVICE, "/tmp/o2d_vtu_bench.vtu", n, d[0], d[1], d[2], d[3], d[4], d[5], 0.5)
print("ok", os.path.getsize("/tmp/o2d_vtu_bench.vtu"))
