"""Generate tests/golden/ref_golden_repr.npz: the reference's own Surfaces::represent_as_particles (src/Surfaces.h:830-873,
through oracle/ref_driver.cpp::o2dref_represent_as_particles) on a body with NON-UNIFORM panel lengths -- an unevenly
sampled ellipse whose longest panels shed several points -- for the two calls the reference makes: panels_affect_panels
(offset 1e-4, ips -1: "use the first panel's length") and the diffusion step's (offset, ips) form.

Run in the build container, where /root/reference exists:  python tests/golden/make_golden_repr.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def body(npan=61):
    t = np.linspace(0, 1, npan, endpoint=False)
    th = -2 * np.pi * (t + 0.35 * np.sin(2 * np.pi * t) / (2 * np.pi) * 2.0)
    nx = (0.6 * np.cos(th)).astype(np.float32); ny = (0.3 * np.sin(th)).astype(np.float32)
    idx = np.empty(2 * npan, np.uint32); idx[0::2] = np.arange(npan); idx[1::2] = (np.arange(npan) + 1) % npan
    vs = (np.sin(3 * th) + 0.2).astype(np.float32)
    return nx, ny, idx, vs


def main():
    assert oracle.build_ref()
    R = oracle.ref("v2")
    nx, ny, idx, vs = body()
    npan = vs.size
    f = R.lib.o2dref_represent_as_particles
    fp = C.POINTER(C.c_float); up = C.POINTER(C.c_uint32)
    f.argtypes = [C.c_int, C.c_int, fp, fp, up, fp, C.c_float, C.c_float, C.c_int, fp, fp, fp]
    out = dict(nx=nx, ny=ny, idx=idx, vs=vs)
    for tag, off, ips in (("pap", 1e-4, -1.0), ("diff", 0.01, 0.02)):
        cap = 4096
        px = np.zeros(cap, np.float32); py = np.zeros(cap, np.float32); pv = np.zeros(cap, np.float32)
        n = f(npan, npan, nx.ctypes.data_as(fp), ny.ctypes.data_as(fp), idx.ctypes.data_as(up), vs.ctypes.data_as(fp), off, ips, cap,
              px.ctypes.data_as(fp), py.ctypes.data_as(fp), pv.ctypes.data_as(fp))
        assert n > npan, "the body must have panels long enough to shed more than one point"
        out[tag + "_args"] = np.array([off, ips], np.float32)
        out[tag + "_x"] = px[:n].copy(); out[tag + "_y"] = py[:n].copy(); out[tag + "_val"] = pv[:n].copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_golden_repr.npz"), **out)


if __name__ == "__main__":
    main()
