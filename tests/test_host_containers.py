"""CPU: the host mirror's containers against the reference's own (no GPU involved)."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_represent_as_particles_on_non_uniform_panels(tmp_path):
    """Surfaces::represent_as_particles of the mirror (omega2d_b200/host/Surfaces.h) on a body whose panel lengths vary
    1 : 6 -- so that long panels shed several points, which panels_affect_panels then maps back by index -- equals the
    reference's own output (tests/golden/ref_golden_repr.npz) bit for bit: same count, positions, strengths."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_repr.npz"))
    exe = os.path.join(ROOT, "build", "mirror_repr")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "omega2d_b200", "host"),
                    os.path.join(ROOT, "tests", "cpp", "mirror_repr.cpp"), "-o", exe], check=True)
    body = tmp_path / "body.txt"
    with open(body, "w") as f:
        f.write(f"{g['vs'].size}\n")
        for x, y, v in zip(g["nx"], g["ny"], g["vs"]):
            f.write(f"{float(x):.9g} {float(y):.9g} {float(v):.9g}\n")
    for tag in ("pap", "diff"):
        off, ips = (float(a) for a in g[tag + "_args"])
        out = subprocess.run([exe, str(body), f"{off:.9g}", f"{ips:.9g}"], check=True, capture_output=True, text=True).stdout.split()
        n = int(out[0])
        vals = np.array(out[1:], np.float32).reshape(n, 3)
        assert n == g[tag + "_x"].size and n > g["vs"].size
        np.testing.assert_array_equal(vals[:, 0], g[tag + "_x"])
        np.testing.assert_array_equal(vals[:, 1], g[tag + "_y"])
        np.testing.assert_array_equal(vals[:, 2], g[tag + "_val"])
