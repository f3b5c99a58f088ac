"""omega2d_b200 -- B200-native direct-sum Biot-Savart evaluation for Omega2D.

The product is the C-ABI shared library ``libomega2d_b200.so`` (sources in
``omega2d_b200/csrc``, ABI in ``include/omega2d_b200.h``) plus the C++ header mirror of the
reference's influence interface in ``omega2d_b200/host``.  This Python package is harness
glue only: a ctypes binding used by ``tests/``, ``bench.py`` and ``__graft_entry__.py``.

There is no CPU fallback anywhere in this package: if the library is missing, or no sm_100
GPU is present, creating a :class:`Context` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# O2D_LIB: load another build of the same library (A/B timing of two builds; development only)
LIB_PATH = os.environ.get("O2D_LIB") or os.path.join(HERE, "libomega2d_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "omega2d_b200.h")

CORES = ("rm", "exp", "wl", "v2", "v3")
CORE_ID = {n: i for i, n in enumerate(CORES)}
# reference results_t (reference src/ResultsType.h:25-31)
VELONLY, VELANDGRAD, PSIONLY, VELANDVORT, VELANDSHEAR = 1, 2, 3, 4, 5
MEM_HOST, MEM_DEVICE = 0, 1
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_COMM = 0, -1, -2, -3, -4, -5

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)
# o2d_bem_solve_fn (include/omega2d_b200.h): int solve(user, stage, time, np, pu, pv, vs, ss)
BEM_SOLVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, _fp, _fp, _fp, _fp)


class PointsView(C.Structure):      # o2d_points_view
    _fields_ = [("n", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p), ("s", C.c_void_p), ("r", C.c_void_p)]


class PanelsView(C.Structure):      # o2d_panels_view
    _fields_ = [("np", C.c_int), ("nnodes", C.c_int), ("nx", C.c_void_p), ("ny", C.c_void_p), ("idx", C.c_void_p),
                ("vs", C.c_void_p), ("ss", C.c_void_p)]


def bem_solve_callback(fn):
    """Wrap fn(stage, time, pu, pv) -> (vs, ss|None) (numpy float32 arrays) as an o2d_bem_solve_fn."""
    def trampoline(user, stage, time, npan, pu, pv, vs, ss):
        try:
            a = np.ctypeslib.as_array(pu, shape=(npan,)).copy()
            b = np.ctypeslib.as_array(pv, shape=(npan,)).copy()
            nvs, nss = fn(int(stage), float(time), a, b)
            np.ctypeslib.as_array(vs, shape=(npan,))[:] = np.asarray(nvs, np.float32)
            if ss and nss is not None:
                np.ctypeslib.as_array(ss, shape=(npan,))[:] = np.asarray(nss, np.float32)
            return 0
        except Exception:   # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1
    return BEM_SOLVE_FN(trampoline)


class O2DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"omega2d_b200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile libomega2d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j", str(max(2, os.cpu_count() or 2))]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libomega2d_b200.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


_lib = None


def load_library() -> C.CDLL:
    """dlopen the product library.  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(omega2d_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.o2d_last_error.restype = C.c_char_p
    lib.o2d_create.argtypes = [C.POINTER(vp), i]
    lib.o2d_destroy.argtypes = [vp]
    lib.o2d_set_stream.argtypes = [vp, vp]
    lib.o2d_synchronize.argtypes = [vp]
    pp = [vp, i, i, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.o2d_points_on_points.argtypes = pp
    lib.o2d_points_on_points_sharded.argtypes = [vp, i, i, i, i, i, vp, vp, vp, vp, i, vp, vp, vp,
                                                 vp, vp, vp, vp, vp, vp, vp]
    lib.o2d_panels_on_points.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp]
    lib.o2d_points_on_panels.argtypes = [vp, i, i, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.o2d_points_vorticity.argtypes = [vp, i, i, vp, vp, vp, vp, i, vp, vp, vp]
    lib.o2d_zero_vels.argtypes = [vp, i, i, vp, vp, vp, vp]
    lib.o2d_finalize_vels.argtypes = [vp, i, i, vp, vp, vp, vp, d, d]
    lib.o2d_find_vels.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp,
                                  i, vp, vp, vp, vp, vp, vp, d, d]
    lib.o2d_advect.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, d, d, d]
    lib.o2d_advect_steps.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, d, d, d, i]
    lib.o2d_find_vels_multi.argtypes = [vp, i, i, i, i, C.POINTER(PointsView), i, C.POINTER(PanelsView),
                                        i, vp, vp, vp, vp, vp, vp, d, d]
    lib.o2d_advect_bodies.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp,
                                      d, d, d, d, C.c_float, C.c_float, i, vp, vp, vp, vp, BEM_SOLVE_FN, vp]
    lib.o2d_panels_on_panels_coeff.argtypes = [vp, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp,
                                               i, i, C.c_longlong, vp]
    lib.o2d_clear_inner_layer.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, i, vp, vp, C.c_float, C.c_float,
                                          C.POINTER(C.c_longlong)]
    lib.o2d_reflect_interior.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, i, vp, vp, C.POINTER(C.c_longlong)]
    lib.o2d_find_vels_sharded.argtypes = [vp, i, i, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp,
                                          i, vp, vp, vp, vp, vp, vp, d, d]
    lib.o2d_points_on_panels_sharded.argtypes = [vp, i, i, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.o2d_points_status.argtypes = [vp, i, i, vp, vp, vp] + [C.POINTER(C.c_float)] * 4
    lib.o2d_add_rot_strengths.argtypes = [vp, i, i, i, vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, vp, vp]
    lib.o2d_vtu_points_size.argtypes = [i, i, i, i, i, C.POINTER(C.c_size_t)]
    lib.o2d_vtu_points.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, d, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.o2d_write_vtu_points.argtypes = [vp, i, C.c_char_p, i, vp, vp, vp, vp, vp, vp, vp, vp, d]
    lib.o2d_group_create.argtypes = [C.POINTER(vp), i, C.POINTER(i)]
    lib.o2d_group_destroy.argtypes = [vp]
    lib.o2d_group_size.argtypes = [vp]
    lib.o2d_group_context.argtypes = [vp, i]
    lib.o2d_group_context.restype = vp
    lib.o2d_group_set_min_pairs.argtypes = [vp, C.c_longlong]
    lib.o2d_group_points_on_points.argtypes = [vp, i, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.o2d_group_find_vels.argtypes = [vp, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, d, d]
    lib.o2d_group_panels_on_points.argtypes = [vp, i, i, vp, vp, vp, vp, vp, i, vp, vp, vp, vp]
    lib.o2d_group_points_on_panels.argtypes = [vp, i, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.external_vel_solver_f_.restype = C.c_float
    lib.external_vel_solver_f_.argtypes = [C.POINTER(i), vp, vp, vp, vp, C.POINTER(i), vp, vp, vp, vp]
    lib.external_vel_solver_tr_f_.restype = C.c_float
    lib.external_vel_solver_tr_f_.argtypes = [C.POINTER(i), vp, vp, vp, vp, C.POINTER(i), vp, vp, vp, vp, vp]
    lib.o2d_set_default_core.argtypes = [i]
    lib.o2d_comm_unique_id.argtypes = [vp]
    lib.o2d_comm_init.argtypes = [vp, i, i, vp]
    lib.o2d_comm_destroy.argtypes = [vp]
    lib.o2d_flops_per_pair.argtypes = [i, i, i, i]
    lib.o2d_enable_timing.argtypes = [vp, i]
    lib.o2d_reset_counters.argtypes = [vp]
    lib.o2d_get_counters.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(d)]
    lib.o2d_device_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(d)]
    lib.o2d_measure_fp32_peak.argtypes = [vp, i, C.POINTER(d)]
    _lib = lib
    return lib


def flops_per_pair(kind: int, core: str = "rm", results: int = VELONLY, blob_targets: bool = True) -> int:
    return int(load_library().o2d_flops_per_pair(kind, CORE_ID[core], results, int(blob_targets)))


class Group:
    """Several GPUs driven from this one process (o2d_group_*): host arrays in, targets cut into one block per device."""

    def __init__(self, devices=None, min_pairs=None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.o2d_group_create(C.byref(h), 0, None)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.o2d_group_create(C.byref(h), len(devices), arr)
        if rc != OK:
            raise O2DError(rc, self.lib.o2d_last_error().decode())
        self._h = h
        if min_pairs is not None:
            self._check(self.lib.o2d_group_set_min_pairs(self._h, int(min_pairs)))

    def _check(self, rc):
        if rc != OK:
            raise O2DError(rc, self.lib.o2d_last_error().decode())

    @property
    def size(self):
        return int(self.lib.o2d_group_size(self._h))

    def close(self):
        if self._h:
            self.lib.o2d_group_destroy(self._h)
            self._h = None

    def points_on_points(self, core, results, sx, sy, ss, sr, tx, ty, tr, tu, tv, tq=None, tux=None, tvx=None, tuy=None, tvy=None):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        self._check(self.lib.o2d_group_points_on_points(self._h, core, int(results), int(sx.size), _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr),
                                                        int(tx.size), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq),
                                                        _ptr(tux), _ptr(tvx), _ptr(tuy), _ptr(tvy)))

    def find_vels(self, core, results, vort, panels, tx, ty, tr, tu, tv, tq=None, fs=(0.0, 0.0)):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        vx, vy, vs, vr = vort
        if panels is not None:
            nx, ny, idx, pvs, pss = panels
            npan, nn = pvs.size, nx.size
        else:
            nx = ny = idx = pvs = pss = None
            npan = nn = 0
        self._check(self.lib.o2d_group_find_vels(self._h, core, int(results), int(vx.size), _ptr(vx), _ptr(vy), _ptr(vs), _ptr(vr),
                                                 int(npan), int(nn), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs), _ptr(pss),
                                                 int(tx.size), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq),
                                                 float(fs[0]), float(fs[1])))

    def panels_on_points(self, panels, tx, ty, tu, tv):
        nx, ny, idx, pvs, pss = panels
        self._check(self.lib.o2d_group_panels_on_points(self._h, int(pvs.size), int(nx.size), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs),
                                                        _ptr(pss), int(tx.size), _ptr(tx), _ptr(ty), _ptr(tu), _ptr(tv)))

    def points_on_panels(self, sx, sy, ss, nx, ny, idx, area, pu, pv):
        self._check(self.lib.o2d_group_points_on_panels(self._h, int(sx.size), _ptr(sx), _ptr(sy), _ptr(ss), int(area.size), int(nx.size),
                                                        _ptr(nx), _ptr(ny), _ptr(idx), _ptr(area), _ptr(pu), _ptr(pv)))


def vtu_points_size(n: int, has_s=True, has_r=True, has_w=False, has_e=False) -> int:
    lib = load_library()
    out = C.c_size_t(0)
    rc = lib.o2d_vtu_points_size(int(n), int(bool(has_s)), int(bool(has_r)), int(bool(has_w)), int(bool(has_e)), C.byref(out))
    if rc != OK:
        raise O2DError(rc, lib.o2d_last_error().decode())
    return int(out.value)


def _ptr(a):
    """numpy array / torch tensor / int address / None -> integer address (or None)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous, "arrays must be contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):   # torch tensor (host or device)
        assert a.is_contiguous(), "tensors must be contiguous"
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)}")


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class Context:
    """One library context = one GPU.  Methods mirror the C ABI one to one; the ``*_np``
    helpers take/return numpy arrays through the O2D_MEM_HOST path (what an unmodified
    Omega2D would use: caller-owned host std::vector storage)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.o2d_create(C.byref(self._h), int(device))
        if rc != OK:
            raise O2DError(rc, self.lib.o2d_last_error().decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.o2d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise O2DError(rc, self.lib.o2d_last_error().decode())

    # -- plumbing
    def set_stream(self, stream_handle):
        self._check(self.lib.o2d_set_stream(self._h, stream_handle))

    def synchronize(self):
        self._check(self.lib.o2d_synchronize(self._h))

    def enable_timing(self, on=True):
        self._check(self.lib.o2d_enable_timing(self._h, int(on)))

    def reset_counters(self):
        self._check(self.lib.o2d_reset_counters(self._h))

    def counters(self):
        a, b, ms = C.c_longlong(), C.c_longlong(), C.c_double()
        self._check(self.lib.o2d_get_counters(self._h, C.byref(a), C.byref(b), C.byref(ms)))
        return {"launches": a.value, "pair_launches": b.value, "pair_ms": ms.value}

    def device_info(self):
        sm, mhz, pk = C.c_int(), C.c_int(), C.c_double()
        self._check(self.lib.o2d_device_info(self._h, C.byref(sm), C.byref(mhz), C.byref(pk)))
        return {"sm_count": sm.value, "sm_clock_mhz": mhz.value, "fp32_peak_tflops": pk.value}

    def measure_fp32_peak(self, packed: bool) -> float:
        t = C.c_double()
        self._check(self.lib.o2d_measure_fp32_peak(self._h, int(packed), C.byref(t)))
        return t.value

    # -- raw C-ABI calls (pointers may be numpy arrays, torch tensors or integer addresses)
    def points_on_points(self, core, results, mem, ns, sx, sy, ss, sr, nt, tx, ty, tr, tu, tv,
                         tq=None, tux=None, tvx=None, tuy=None, tvy=None):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        self._check(self.lib.o2d_points_on_points(
            self._h, core, int(results), int(mem), int(ns), _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr),
            int(nt), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq),
            _ptr(tux), _ptr(tvx), _ptr(tuy), _ptr(tvy)))

    def points_on_points_sharded(self, core, results, mem, ns_local, ns_block, sx, sy, ss, sr, nt,
                                 tx, ty, tr, tu, tv, tq=None, tux=None, tvx=None, tuy=None, tvy=None):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        self._check(self.lib.o2d_points_on_points_sharded(
            self._h, core, int(results), int(mem), int(ns_local), int(ns_block),
            _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr), int(nt), _ptr(tx), _ptr(ty), _ptr(tr),
            _ptr(tu), _ptr(tv), _ptr(tq), _ptr(tux), _ptr(tvx), _ptr(tuy), _ptr(tvy)))

    def panels_on_points(self, mem, np_, nnodes, nx, ny, idx, vs, ss, nt, tx, ty, tu, tv):
        self._check(self.lib.o2d_panels_on_points(
            self._h, int(mem), int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(vs), _ptr(ss),
            int(nt), _ptr(tx), _ptr(ty), _ptr(tu), _ptr(tv)))

    def points_on_panels(self, mem, ns, sx, sy, ss, np_, nnodes, nx, ny, idx, area, pu, pv):
        self._check(self.lib.o2d_points_on_panels(
            self._h, int(mem), int(ns), _ptr(sx), _ptr(sy), _ptr(ss), int(np_), int(nnodes),
            _ptr(nx), _ptr(ny), _ptr(idx), _ptr(area), _ptr(pu), _ptr(pv)))

    def points_vorticity(self, mem, ns, sx, sy, ss, sr, nt, tx, ty, tw):
        self._check(self.lib.o2d_points_vorticity(
            self._h, int(mem), int(ns), _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr), int(nt),
            _ptr(tx), _ptr(ty), _ptr(tw)))

    def zero_vels(self, mem, n, u, v, w=None, e=None):
        self._check(self.lib.o2d_zero_vels(self._h, int(mem), int(n), _ptr(u), _ptr(v), _ptr(w), _ptr(e)))

    def finalize_vels(self, mem, n, u, v, w=None, e=None, fs=(0.0, 0.0)):
        self._check(self.lib.o2d_finalize_vels(self._h, int(mem), int(n), _ptr(u), _ptr(v), _ptr(w), _ptr(e),
                                               float(fs[0]), float(fs[1])))

    def find_vels(self, core, results, mem, nv, vx, vy, vs, vr, np_, nnodes, nx, ny, idx, pvs, pss,
                  nt, tx, ty, tr, tu, tv, tq=None, fs=(0.0, 0.0)):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        self._check(self.lib.o2d_find_vels(
            self._h, core, int(results), int(mem), int(nv), _ptr(vx), _ptr(vy), _ptr(vs), _ptr(vr),
            int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs), _ptr(pss),
            int(nt), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq), float(fs[0]), float(fs[1])))

    def advect(self, core, order, mem, n, x, y, s, r, u, v, np_=0, nnodes=0, nx=None, ny=None, idx=None,
               pvs=None, pss=None, fs=(0.0, 0.0), dt=0.0, nsteps=None):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        if nsteps is not None:
            self._check(self.lib.o2d_advect_steps(
                self._h, core, int(order), int(mem), int(n), _ptr(x), _ptr(y), _ptr(s), _ptr(r), _ptr(u), _ptr(v),
                int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs), _ptr(pss),
                float(fs[0]), float(fs[1]), float(dt), int(nsteps)))
            return
        self._check(self.lib.o2d_advect(
            self._h, core, int(order), int(mem), int(n), _ptr(x), _ptr(y), _ptr(s), _ptr(r), _ptr(u), _ptr(v),
            int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs), _ptr(pss),
            float(fs[0]), float(fs[1]), float(dt)))

    def panels_on_panels_coeff(self, mem, nsrc, nsnodes, snx, sny, sidx, sarea, ntarg, ntnodes, tnx, tny, tidx,
                               tarea, ttx, tty, tnnx, tnny, both, same, ld, coeffs):
        self._check(self.lib.o2d_panels_on_panels_coeff(
            self._h, int(mem), int(nsrc), int(nsnodes), _ptr(snx), _ptr(sny), _ptr(sidx), _ptr(sarea),
            int(ntarg), int(ntnodes), _ptr(tnx), _ptr(tny), _ptr(tidx), _ptr(tarea), _ptr(ttx), _ptr(tty),
            _ptr(tnnx), _ptr(tnny), int(both), int(same), int(ld), _ptr(coeffs)))

    def clear_inner_layer(self, method, mem, np_, nnodes, nx, ny, idx, snx, sny, n, tx, ty, cutoff_mult, ips,
                          want_count=True):
        moved = C.c_longlong(0)
        self._check(self.lib.o2d_clear_inner_layer(
            self._h, int(method), int(mem), int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(snx), _ptr(sny),
            int(n), _ptr(tx), _ptr(ty), float(cutoff_mult), float(ips), C.byref(moved) if want_count else None))
        return moved.value

    def reflect_interior(self, mem, np_, nnodes, nx, ny, idx, snx, sny, n, tx, ty, want_count=True):
        moved = C.c_longlong(0)
        self._check(self.lib.o2d_reflect_interior(
            self._h, int(mem), int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(snx), _ptr(sny),
            int(n), _ptr(tx), _ptr(ty), C.byref(moved) if want_count else None))
        return moved.value

    def find_vels_sharded(self, core, results, mem, nv_local, nv_block, vx, vy, vs, vr, np_, nnodes, nx, ny, idx,
                          pvs, pss, nt, tx, ty, tr, tu, tv, tq=None, fs=(0.0, 0.0)):
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        self._check(self.lib.o2d_find_vels_sharded(
            self._h, core, int(results), int(mem), int(nv_local), int(nv_block), _ptr(vx), _ptr(vy), _ptr(vs), _ptr(vr),
            int(np_), int(nnodes), _ptr(nx), _ptr(ny), _ptr(idx), _ptr(pvs), _ptr(pss),
            int(nt), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq), float(fs[0]), float(fs[1])))

    def points_on_panels_sharded(self, mem, ns_local, sx, sy, ss, np_, nnodes, nx, ny, idx, area, pu, pv):
        self._check(self.lib.o2d_points_on_panels_sharded(
            self._h, int(mem), int(ns_local), _ptr(sx), _ptr(sy), _ptr(ss), int(np_), int(nnodes),
            _ptr(nx), _ptr(ny), _ptr(idx), _ptr(area), _ptr(pu), _ptr(pv)))

    def points_status(self, mem, n, x, y, s):
        """-> (total_circ, impulse_x, impulse_y, max_str) as Python floats (float32 values)."""
        out = [C.c_float() for _ in range(4)]
        self._check(self.lib.o2d_points_status(self._h, int(mem), int(n), _ptr(x), _ptr(y), _ptr(s),
                                               *[C.byref(o) for o in out]))
        return tuple(o.value for o in out)

    def vtu_points(self, mem, n, x, y, s, r, u, v, time, w=None, e=None, cap=None, out=None):
        """The .vtu file image of Points::write_vtk as bytes (o2d_vtu_points).  With `out` (a host uint8 buffer, e.g.
        pinned) the image is written there and its size is returned instead."""
        need = vtu_points_size(n, s is not None, r is not None, w is not None, e is not None)
        cap = need if cap is None else int(cap)
        if out is not None:
            wrote = C.c_size_t(0)
            self._check(self.lib.o2d_vtu_points(self._h, int(mem), int(n), _ptr(x), _ptr(y), _ptr(s), _ptr(r), _ptr(w), _ptr(e),
                                                _ptr(u), _ptr(v), float(time), _ptr(out), int(out.numel() if hasattr(out, "numel") else out.size),
                                                C.byref(wrote)))
            return wrote.value
        buf = np.empty(max(cap, 1), dtype=np.uint8)
        wrote = C.c_size_t(0)
        self._check(self.lib.o2d_vtu_points(self._h, int(mem), int(n), _ptr(x), _ptr(y), _ptr(s), _ptr(r), _ptr(w), _ptr(e),
                                            _ptr(u), _ptr(v), float(time), _ptr(buf), cap, C.byref(wrote)))
        assert wrote.value == need
        return buf[:wrote.value].tobytes()

    def write_vtu_points(self, mem, path, n, x, y, s, r, u, v, time, w=None, e=None):
        self._check(self.lib.o2d_write_vtu_points(self._h, int(mem), str(path).encode(), int(n), _ptr(x), _ptr(y), _ptr(s),
                                                  _ptr(r), _ptr(w), _ptr(e), _ptr(u), _ptr(v), float(time)))

    def add_rot_strengths(self, mem, np_, nnodes, ux, uy, idx, area, center, rotvel, vs, ss):
        self._check(self.lib.o2d_add_rot_strengths(self._h, int(mem), int(np_), int(nnodes), _ptr(ux), _ptr(uy), _ptr(idx),
                                                   _ptr(area), float(center[0]), float(center[1]), float(rotvel),
                                                   _ptr(vs), _ptr(ss)))

    # -- comm
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.o2d_comm_unique_id(buf)
        if rc != OK:
            raise O2DError(rc, lib.o2d_last_error().decode())
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        assert len(uid) == 128
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.o2d_comm_init(self._h, int(nranks), int(rank), buf))

    # -- numpy conveniences over the host-memory path (same argument order as the oracle wrappers)
    def points_on_points_np(self, core, results, sx, sy, ss, sr, tx, ty, tr=None, init=None):
        sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
        nt = tx.size
        nout = {VELONLY: 2, VELANDVORT: 3, VELANDSHEAR: 3, VELANDGRAD: 6}.get(results, 2)
        outs = [np.zeros(nt, np.float32) for _ in range(nout)]
        if init is not None:
            for o, i0 in zip(outs, init):
                o[:] = i0
        tq = outs[2] if nout == 3 else None
        g = outs[2:6] if nout == 6 else [None] * 4
        self.points_on_points(core, results, MEM_HOST, sx.size, sx, sy, ss, sr, nt, tx, ty, tr,
                              outs[0], outs[1], tq, *g)
        return tuple(outs)

    def find_vels_np(self, core, results, vort, panels, tx, ty, tr=None, fs=(0.0, 0.0)):
        """vort = (x, y, s, r) or None; panels = (nx, ny, idx, vs, ss|None) or None."""
        tx, ty, tr = map(_f32, (tx, ty, tr))
        nt = tx.size
        vx, vy, vs, vr = map(_f32, vort) if vort is not None else (None,) * 4
        if panels is not None:
            nx, ny, pvs, pss = map(_f32, (panels[0], panels[1], panels[3], panels[4]))
            idx = np.ascontiguousarray(panels[2], dtype=np.uint32)
            npan, nn = pvs.size, nx.size
        else:
            nx = ny = idx = pvs = pss = None
            npan = nn = 0
        tu, tv = np.empty(nt, np.float32), np.empty(nt, np.float32)
        tq = np.empty(nt, np.float32) if results in (VELANDVORT, VELANDSHEAR) else None
        self.find_vels(core, results, MEM_HOST, 0 if vx is None else vx.size, vx, vy, vs, vr, npan, nn, nx, ny, idx,
                       pvs, pss, nt, tx, ty, tr, tu, tv, tq, fs)
        return (tu, tv) if tq is None else (tu, tv, tq)

    def find_vels_multi_np(self, core, results, vorts, panels, tx, ty, tr=None, fs=(0.0, 0.0)):
        """vorts = list of (x, y, s, r); panels = list of (nx, ny, idx, vs, ss|None): o2d_find_vels_multi on host arrays."""
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        tx, ty, tr = map(_f32, (tx, ty, tr))
        nt = tx.size
        keep = []
        pv = (PointsView * max(len(vorts), 1))()
        for k, vv in enumerate(vorts):
            a = [_f32(q) for q in vv]
            keep.append(a)
            pv[k] = PointsView(a[0].size, *[q.ctypes.data for q in a])
        qv = (PanelsView * max(len(panels), 1))()
        for k, pp in enumerate(panels):
            nx, ny, pvs = map(_f32, (pp[0], pp[1], pp[3]))
            pss = None if pp[4] is None else _f32(pp[4])
            idx = np.ascontiguousarray(pp[2], dtype=np.uint32)
            keep.append((nx, ny, idx, pvs, pss))
            qv[k] = PanelsView(pvs.size, nx.size, nx.ctypes.data, ny.ctypes.data, idx.ctypes.data, pvs.ctypes.data,
                               None if pss is None else pss.ctypes.data)
        tu, tv = np.empty(nt, np.float32), np.empty(nt, np.float32)
        tq = np.empty(nt, np.float32) if results in (VELANDVORT, VELANDSHEAR) else None
        self._check(self.lib.o2d_find_vels_multi(self._h, core, int(results), MEM_HOST, len(vorts), pv, len(panels), qv,
                                                 nt, _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv), _ptr(tq),
                                                 float(fs[0]), float(fs[1])))
        return (tu, tv) if tq is None else (tu, tv, tq)

    def advect_np(self, core, order, x, y, s, r, dt, panels=None, fs=(0.0, 0.0)):
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r = map(_f32, (s, r))
        if panels is not None:
            nx, ny, pvs, pss = map(_f32, (panels[0], panels[1], panels[3], panels[4]))
            idx = np.ascontiguousarray(panels[2], dtype=np.uint32)
            npan, nn = pvs.size, nx.size
        else:
            nx = ny = idx = pvs = pss = None
            npan = nn = 0
        u, v = np.empty(x.size, np.float32), np.empty(x.size, np.float32)
        self.advect(core, order, MEM_HOST, x.size, x, y, s, r, u, v, npan, nn, nx, ny, idx, pvs, pss, fs, dt)
        return x, y, u, v

    def advect_bodies_np(self, core, order, x, y, s, r, panels, area, normals, solve, dt, ips, fs=(0.0, 0.0), time=0.0,
                         cutoff_mult=None, fldpt=None):
        """One convection step with bodies (o2d_advect_bodies).  panels = (nx, ny, idx, vs, ss|None); normals = (snx, sny);
        solve(stage, time, pu, pv) -> (vs, ss); fldpt = (fx, fy) inert tracers or None.  Returns (x, y, u, v, vs, ss) after
        the step, followed by (fx, fy, fu, fv) when tracers were given."""
        core = CORE_ID[core] if isinstance(core, str) else int(core)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r, area = map(_f32, (s, r, area))
        nx, ny = map(_f32, panels[:2])
        idx = np.ascontiguousarray(panels[2], dtype=np.uint32)
        pvs = np.array(panels[3], np.float32)
        pss = None if panels[4] is None else np.array(panels[4], np.float32)
        snx, sny = map(_f32, normals)
        u, v = np.empty(x.size, np.float32), np.empty(x.size, np.float32)
        if cutoff_mult is None:
            cutoff_mult = np.float32(0.5 / np.sqrt(2.0 * np.pi))   # src/Convection.h:306
        cb = bem_solve_callback(solve)
        nf = 0
        fx = fy = fu = fv = None
        if fldpt is not None:
            fx, fy = np.array(fldpt[0], np.float32), np.array(fldpt[1], np.float32)
            nf = fx.size
            fu, fv = np.empty(nf, np.float32), np.empty(nf, np.float32)
        self._check(self.lib.o2d_advect_bodies(
            self._h, core, int(order), MEM_HOST, x.size, _ptr(x), _ptr(y), _ptr(s), _ptr(r), _ptr(u), _ptr(v),
            pvs.size, nx.size, _ptr(nx), _ptr(ny), _ptr(idx), _ptr(area), _ptr(snx), _ptr(sny), _ptr(pvs), _ptr(pss),
            float(fs[0]), float(fs[1]), float(time), float(dt), float(cutoff_mult), float(ips),
            nf, _ptr(fx), _ptr(fy), _ptr(fu), _ptr(fv), cb, None))
        return (x, y, u, v, pvs, pss) if fldpt is None else (x, y, u, v, pvs, pss, fx, fy, fu, fv)

    def panels_on_panels_coeff_np(self, src, sarea, targ, tbases, both=True, same=False, ld=None):
        """src/targ = (nx, ny, idx); tbases = (ttx, tty, tnx, tny, tarea).  Returns (cols, ld) float32."""
        snx, sny = map(_f32, src[:2]); sidx = np.ascontiguousarray(src[2], dtype=np.uint32)
        tnx, tny = map(_f32, targ[:2]); tidx = np.ascontiguousarray(targ[2], dtype=np.uint32)
        ttx, tty, tnnx, tnny, tarea = map(_f32, tbases)
        sarea = _f32(sarea)
        ns, nt = sidx.size // 2, tidx.size // 2
        unk = 2 if both else 1
        ld = unk * nt if ld is None else int(ld)
        c = np.zeros((unk * ns, ld), np.float32)
        self.panels_on_panels_coeff(MEM_HOST, ns, snx.size, snx, sny, sidx, sarea, nt, tnx.size, tnx, tny, tidx, tarea,
                                    ttx, tty, tnnx, tnny, both, same, ld, c)
        return c

    def clear_inner_layer_np(self, nx, ny, idx, snx, sny, x, y, cutoff_mult, ips):
        nx, ny, snx, sny = map(_f32, (nx, ny, snx, sny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        moved = self.clear_inner_layer(1, MEM_HOST, snx.size, nx.size, nx, ny, idx, snx, sny, x.size, x, y, cutoff_mult, ips)
        return x, y, moved

    def reflect_interior_np(self, nx, ny, idx, snx, sny, x, y):
        nx, ny, snx, sny = map(_f32, (nx, ny, snx, sny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        moved = self.reflect_interior(MEM_HOST, snx.size, nx.size, nx, ny, idx, snx, sny, x.size, x, y)
        return x, y, moved

    def panels_on_points_np(self, nx, ny, idx, vs, ss, tx, ty, init=None):
        nx, ny, vs, ss, tx, ty = map(_f32, (nx, ny, vs, ss, tx, ty))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        nt = tx.size
        tu, tv = np.zeros(nt, np.float32), np.zeros(nt, np.float32)
        if init is not None:
            tu[:], tv[:] = init
        self.panels_on_points(MEM_HOST, vs.size, nx.size, nx, ny, idx, vs, ss, nt, tx, ty, tu, tv)
        return tu, tv

    def points_on_panels_np(self, sx, sy, ss, nx, ny, idx, area, init=None):
        sx, sy, ss, nx, ny, area = map(_f32, (sx, sy, ss, nx, ny, area))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        npan = idx.size // 2
        pu, pv = np.zeros(npan, np.float32), np.zeros(npan, np.float32)
        if init is not None:
            pu[:], pv[:] = init
        self.points_on_panels(MEM_HOST, sx.size, sx, sy, ss, npan, nx.size, nx, ny, idx, area, pu, pv)
        return pu, pv

    def points_vorticity_np(self, sx, sy, ss, sr, tx, ty, init=None):
        sx, sy, ss, sr, tx, ty = map(_f32, (sx, sy, ss, sr, tx, ty))
        tw = np.zeros(tx.size, np.float32)
        if init is not None:
            tw[:] = init
        self.points_vorticity(MEM_HOST, sx.size, sx, sy, ss, sr, tx.size, tx, ty, tw)
        return tw


def external_vel_solver(sx, sy, ss, sr, tx, ty, tr=None, tu=None, tv=None):
    """Call the legacy plugin symbols exactly as Omega2D does (reference src/Influence.h:82-88)."""
    lib = load_library()
    sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
    ns, nt = C.c_int(sx.size), C.c_int(tx.size)
    tu = np.zeros(tx.size, np.float32) if tu is None else np.array(tu, np.float32)
    tv = np.zeros(tx.size, np.float32) if tv is None else np.array(tv, np.float32)
    if tr is None:
        flops = lib.external_vel_solver_f_(C.byref(ns), _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr),
                                           C.byref(nt), _ptr(tx), _ptr(ty), _ptr(tu), _ptr(tv))
    else:
        flops = lib.external_vel_solver_tr_f_(C.byref(ns), _ptr(sx), _ptr(sy), _ptr(ss), _ptr(sr),
                                              C.byref(nt), _ptr(tx), _ptr(ty), _ptr(tr), _ptr(tu), _ptr(tv))
    if flops < 0:
        raise O2DError(int(flops), lib.o2d_last_error().decode())
    return tu, tv, float(flops)
