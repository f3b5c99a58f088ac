"""Generates tests/golden/ref_golden_c5.npz with the UNMODIFIED reference compiled here (oracle/_ref): BASELINE.json
configs[4] (C5) source count -- 16 777 216 Winckelmans-Leonard particles -- summed by the reference's own
points_affect_points on 256 strided targets.  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_c5.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from omega2d_b200 import workloads as W

N = 16_777_216
NT = 65_536        # targets the GPU test evaluates (strided through the set)
NCHECK = 256       # of which the reference evaluates these (strided through the NT)


def main():
    sx, sy, ss, sr = W.gaussian_blob(N, seed=2024)
    tsub = W.strided_subset(N, NT)
    csub = W.strided_subset(NT, NCHECK)
    sub = tsub[csub]
    u, v = oracle.ref("wl").points_on_points(oracle.VELONLY, sx, sy, ss, sr, sx[sub], sy[sub], sr[sub])
    out = dict(c5_targets=tsub, c5_check=csub, c5_tx=sx[sub].copy(), c5_ty=sy[sub].copy(), c5_tr=sr[sub].copy(),
               c5_wl=np.stack([u, v]),
               c5_input_checksum=np.array([a.astype(np.float64).sum() for a in (sx, sy, ss, sr)]))
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_c5.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
