"""Generate tests/golden/ref_golden_vtu.npz from the UNMODIFIED reference build (oracle/_ref): the .vtu files the
reference's own Points<float>::write_vtk (src/Points.h:556-740, VtkXmlWriter.h, cppcodec, tinyxml2) writes for small
particle collections -- blobs, inert field points, with and without vorticity / shear-rate arrays, and particle
counts covering every base64 tail (n mod 3 for the Float32/Int32 arrays, n mod 12 for the UInt8 types array).

    python tests/golden/make_golden_vtu.py        (in the build container: needs /root/reference)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402

# (tag, n, blobs?, vorticity?, shear?)
CASES = [("n1", 1, True, False, False), ("n2", 2, True, False, False), ("n3", 3, True, False, False),
         ("n4_inert", 4, False, False, False), ("n5_w", 5, True, True, False), ("n7_inert_w", 7, False, True, False),
         ("n11_we", 11, True, True, True), ("n12_e", 12, True, False, True), ("n13_inert_e", 13, False, False, True),
         ("n250", 250, True, False, False), ("n1001_w", 1001, True, True, False)]


def case_arrays(tag, n):
    rng = np.random.default_rng(sum(map(ord, tag)) + 7 * n)
    a = [rng.standard_normal(n).astype(np.float32) for _ in range(8)]
    # values whose bytes hit every character class of the alphabet, incl. '+' and '/'
    a[0][0] = np.frombuffer(bytes([0xFB, 0xEF, 0xBE, 0xFF]), np.float32)[0]
    a[4][-1] = np.float32(0.0)
    return a


def main():
    assert oracle.build_ref(), "reference build (oracle/_ref) unavailable"
    R = oracle.ref("v2")
    out = {"tags": np.array([c[0] for c in CASES])}
    with tempfile.TemporaryDirectory() as d:
        for k, (tag, n, blobs, hw, he) in enumerate(CASES):
            x, y, s, r, u, v, w, e = case_arrays(tag, n)
            time = 0.125 * (k + 1) + 1e-3
            img, name = R.write_vtk_points(d, k, 10 * k + 3, time, x, y, s if blobs else None, r if blobs else None, u, v,
                                           w=w if hw else None, e=e if he else None)
            assert name == f"{'part' if blobs else 'fldpt'}_{k:02d}_{10 * k + 3:05d}.vtu", name
            out[f"{tag}_in"] = np.stack([x, y, s, r, u, v, w, e])
            out[f"{tag}_flags"] = np.array([blobs, hw, he])
            out[f"{tag}_time"] = np.float64(time)
            out[f"{tag}_file"] = np.frombuffer(img, np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "ref_golden_vtu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
