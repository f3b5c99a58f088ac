import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import omega2d_b200 as o2d, oracle
from omega2d_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = o2d.Context(0)
sx, sy, ss, sr = W.gaussian_blob(n, seed=2024)
dev = torch.device("cuda:0")
t = [torch.from_numpy(a).to(dev) for a in (sx, sy, ss, sr)]
res = []
for rep in range(3):
    tu = torch.zeros(n, device=dev); tv = torch.zeros(n, device=dev)
    ctx.points_on_points("rm", 1, o2d.MEM_DEVICE, n, *t, n, t[0], t[1], t[3], tu, tv)
    ctx.synchronize()
    res.append(tu.cpu().numpy())
for rep in range(2):
    u, v = ctx.points_on_points_np("rm", 1, sx, sy, ss, sr, sx, sy, sr)
    res.append(u)
for i in range(1, len(res)):
    d = res[i] != res[0]
    print("run", i, "mismatches", int(d.sum()), "first idx", np.nonzero(d)[0][:10], "maxabs", float(np.abs(res[i]-res[0]).max()))
sub = W.strided_subset(n, 512)
want = oracle.port().points_on_points("rm", 1, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])
sc = np.abs(want[0]).max()
for i, r in enumerate(res):
    print("run", i, "err vs oracle", float(np.abs(r[sub]-want[0]).max()/sc))
