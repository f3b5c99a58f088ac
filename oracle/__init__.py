"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-ends for the two CPU checkers of the Omega2D Biot-Savart hot path:

* ``port``  -- ``oracle/libo2d_oracle.so``, the plain-C restatement in ``o2d_oracle.c``
  (all five cores selectable at run time);
* ``ref(core)`` -- ``oracle/_ref/libo2dref_<core>.so``, the UNMODIFIED reference templates
  (``points_affect_points<float,double>`` etc., reference src/Influence.h) compiled by
  ``oracle/build_ref.sh``; one library per core because the reference picks the core with a
  ``#define`` (reference src/CoreFunc.h:35-39).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / reference
arm may import this package.  The product (``omega2d_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CORES = ("rm", "exp", "wl", "v2", "v3")
CORE_ID = {name: i for i, name in enumerate(CORES)}
# reference results_t values (reference src/ResultsType.h:25-31)
VELONLY, VELANDGRAD, PSIONLY, VELANDVORT, VELANDSHEAR = 1, 2, 3, 4, 5

_fp = C.POINTER(C.c_float)
_up = C.POINTER(C.c_uint32)


def _f(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_fp)


def _u(a):
    assert a.dtype == np.uint32 and a.flags.c_contiguous
    return a.ctypes.data_as(_up)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def build_port() -> str:
    """(Re)build the C restatement if missing or stale; returns the library path."""
    so = os.path.join(HERE, "libo2d_oracle.so")
    src = os.path.join(HERE, "o2d_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


def build_ref(reference: str = "/root/reference") -> bool:
    """Build oracle/_ref from the reference sources when they are present.  Returns True
    when all five libraries exist afterwards (prebuilt ones count)."""
    if os.path.isdir(os.path.join(reference, "src")):
        drv = os.path.join(HERE, "ref_driver.cpp")
        stale = any(
            not os.path.exists(ref_path(c)) or os.path.getmtime(ref_path(c)) < os.path.getmtime(drv)
            for c in CORES
        )
        # the drop-in binary (reference containers + this repository's influence headers) follows its sources
        dropin = os.path.join(HERE, "_ref", "ref_dropin_v2")
        root = os.path.dirname(HERE)
        seam = os.path.join(HERE, "_ref", "ref_external_seam_v2")
        deps = [os.path.join(root, "tests", "cpp", "ref_dropin.cpp"), os.path.join(root, "tests", "cpp", "ref_external_seam.cpp")] + \
               [os.path.join(root, "omega2d_b200", "host", f) for f in os.listdir(os.path.join(root, "omega2d_b200", "host")) if f.endswith(".h")]
        if os.path.exists(os.path.join(root, "omega2d_b200", "libomega2d_b200.so")):
            stale = stale or not os.path.exists(dropin) or not os.path.exists(seam) or \
                any(min(os.path.getmtime(dropin), os.path.getmtime(seam)) < os.path.getmtime(d) for d in deps)
        if stale:
            env = dict(os.environ, O2D_REFERENCE=reference)
            subprocess.run([os.path.join(HERE, "build_ref.sh")], check=True, env=env)
    return have_ref()


def ref_path(core: str) -> str:
    return os.path.join(HERE, "_ref", f"libo2dref_{core}.so")


def have_ref() -> bool:
    return all(os.path.exists(ref_path(c)) for c in CORES)


class Port:
    """The plain-C restatement (o2d_oracle.c)."""

    kind = "port"

    def __init__(self):
        self.lib = C.CDLL(build_port())

    def flops(self, core: str, which: int, has_tr: bool) -> int:
        return int(self.lib.o2d_oracle_flops(CORE_ID[core], which, int(has_tr)))

    def points_on_points(self, core, results, sx, sy, ss, sr, tx, ty, tr=None, tu=None, tv=None, tq=None):
        sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
        nt = tx.size
        tu = np.zeros(nt, np.float32) if tu is None else np.array(tu, np.float32)
        tv = np.zeros(nt, np.float32) if tv is None else np.array(tv, np.float32)
        want_q = results in (VELANDVORT, VELANDSHEAR)
        if want_q:
            tq = np.zeros(nt, np.float32) if tq is None else np.array(tq, np.float32)
        rc = self.lib.o2d_oracle_points_on_points(
            CORE_ID[core], int(results), int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr),
            int(nt), _f(tx), _f(ty), _f(tr), _f(tu), _f(tv), _f(tq) if want_q else None)
        if rc != 0:
            raise ValueError(f"oracle rejected results type {results}")
        return (tu, tv, tq) if want_q else (tu, tv)

    def points_on_points_grad(self, core, sx, sy, ss, sr, tx, ty, tr=None, skip_self=False):
        sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
        nt = tx.size
        out = [np.zeros(nt, np.float32) for _ in range(6)]
        rc = self.lib.o2d_oracle_points_on_points_grad(
            CORE_ID[core], int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr),
            int(nt), _f(tx), _f(ty), _f(tr), *[_f(o) for o in out], int(skip_self))
        assert rc == 0
        return tuple(out)  # u, v, ux, vx, uy, vy

    def panels_on_points(self, nx, ny, idx, vs, ss, tx, ty, tu=None, tv=None):
        nx, ny, vs, ss, tx, ty = map(_f32, (nx, ny, vs, ss, tx, ty))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        nt = tx.size
        tu = np.zeros(nt, np.float32) if tu is None else np.array(tu, np.float32)
        tv = np.zeros(nt, np.float32) if tv is None else np.array(tv, np.float32)
        rc = self.lib.o2d_oracle_panels_on_points(
            int(vs.size), int(nx.size), _f(nx), _f(ny), _u(idx), _f(vs), _f(ss),
            int(nt), _f(tx), _f(ty), _f(tu), _f(tv))
        assert rc == 0
        return tu, tv

    def panel_areas(self, nx, ny, idx):
        nx, ny = map(_f32, (nx, ny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        area = np.zeros(idx.size // 2, np.float32)
        self.lib.o2d_oracle_panel_areas(int(area.size), _f(nx), _f(ny), _u(idx), _f(area))
        return area

    def points_on_panels(self, sx, sy, ss, nx, ny, idx, area=None, pu=None, pv=None):
        sx, sy, ss, nx, ny = map(_f32, (sx, sy, ss, nx, ny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        npan = idx.size // 2
        area = self.panel_areas(nx, ny, idx) if area is None else _f32(area)
        pu = np.zeros(npan, np.float32) if pu is None else np.array(pu, np.float32)
        pv = np.zeros(npan, np.float32) if pv is None else np.array(pv, np.float32)
        rc = self.lib.o2d_oracle_points_on_panels(
            int(sx.size), _f(sx), _f(sy), _f(ss), int(npan), int(nx.size), _f(nx), _f(ny), _u(idx),
            _f(area), _f(pu), _f(pv))
        assert rc == 0
        return pu, pv

    def points_vorticity(self, sx, sy, ss, sr, tx, ty, tw=None):
        sx, sy, ss, sr, tx, ty = map(_f32, (sx, sy, ss, sr, tx, ty))
        nt = tx.size
        tw = np.zeros(nt, np.float32) if tw is None else np.array(tw, np.float32)
        rc = self.lib.o2d_oracle_points_vorticity(
            int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr), int(nt), _f(tx), _f(ty), _f(tw))
        assert rc == 0
        return tw

    def finalize_vels(self, u, v, w=None, e=None, fs=(0.0, 0.0)):
        u, v = np.array(u, np.float32), np.array(v, np.float32)
        w = None if w is None else np.array(w, np.float32)
        e = None if e is None else np.array(e, np.float32)
        self.lib.o2d_oracle_finalize_vels.argtypes = [C.c_int, _fp, _fp, _fp, _fp, C.c_double, C.c_double]
        self.lib.o2d_oracle_finalize_vels(int(u.size), _f(u), _f(v), _f(w), _f(e), float(fs[0]), float(fs[1]))
        return u, v, w, e

    def _panels(self, panels):
        if panels is None:
            return 0, 0, None, None, None, None, None
        nx, ny, pvs, pss = map(_f32, (panels[0], panels[1], panels[3], panels[4]))
        idx = np.ascontiguousarray(panels[2], dtype=np.uint32)
        return int(pvs.size), int(nx.size), nx, ny, idx, pvs, pss

    def find_vels(self, core, results, vort, panels, tx, ty, tr=None, fs=(0.0, 0.0)):
        """Convection::find_vels for one particle + one surface collection on one target set."""
        tx, ty, tr = map(_f32, (tx, ty, tr))
        nt = tx.size
        vx, vy, vs, vr = map(_f32, vort) if vort is not None else (None,) * 4
        npan, nn, nx, ny, idx, pvs, pss = self._panels(panels)
        tu, tv = np.zeros(nt, np.float32), np.zeros(nt, np.float32)
        tq = np.zeros(nt, np.float32) if results in (VELANDVORT, VELANDSHEAR) else None
        f = self.lib.o2d_oracle_find_vels
        f.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _up, _fp, _fp,
                      C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_double, C.c_double]
        rc = f(CORE_ID[core], int(results), 0 if vx is None else int(vx.size), _f(vx), _f(vy), _f(vs), _f(vr),
               npan, nn, _f(nx), _f(ny), None if idx is None else _u(idx), _f(pvs), _f(pss),
               int(nt), _f(tx), _f(ty), _f(tr), _f(tu), _f(tv), _f(tq), float(fs[0]), float(fs[1]))
        assert rc == 0
        return (tu, tv) if tq is None else (tu, tv, tq)

    def advect(self, core, order, x, y, s, r, dt, panels=None, fs=(0.0, 0.0)):
        """One convection step (Euler / Ralston RK2 / RK3) of a particle collection; returns x, y, u, v."""
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r = map(_f32, (s, r))
        npan, nn, nx, ny, idx, pvs, pss = self._panels(panels)
        u, v = np.zeros(x.size, np.float32), np.zeros(x.size, np.float32)
        f = self.lib.o2d_oracle_advect
        f.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _up,
                      _fp, _fp, C.c_double, C.c_double, C.c_double]
        rc = f(CORE_ID[core], int(order), int(x.size), _f(x), _f(y), _f(s), _f(r), _f(u), _f(v),
               npan, nn, _f(nx), _f(ny), None if idx is None else _u(idx), _f(pvs), _f(pss),
               float(fs[0]), float(fs[1]), float(dt))
        assert rc == 0
        return x, y, u, v
    def panel_bases(self, nx, ny, idx):
        """Surfaces::compute_bases: returns ttx, tty, tnx, tny, area."""
        nx, ny = map(_f32, (nx, ny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        n = idx.size // 2
        out = [np.zeros(n, np.float32) for _ in range(5)]
        self.lib.o2d_oracle_panel_bases(int(n), _f(nx), _f(ny), _u(idx), *[_f(o) for o in out])
        return tuple(out)

    def panels_on_panels_coeff(self, src, targ=None, both=True, ld=None):
        """src/targ = (nx, ny, idx); targ None = same surface.  Returns the column-major block as a
        (cols, ld) array (row index fastest), like the reference's Vector<S>."""
        same = targ is None
        targ = src if same else targ
        snx, sny = map(_f32, src[:2]); sidx = np.ascontiguousarray(src[2], dtype=np.uint32)
        tnx, tny = map(_f32, targ[:2]); tidx = np.ascontiguousarray(targ[2], dtype=np.uint32)
        ns, nt = sidx.size // 2, tidx.size // 2
        ttx, tty, tnx_, tny_, ta = self.panel_bases(tnx, tny, tidx)
        sa = self.panel_bases(snx, sny, sidx)[4]
        unk = 2 if both else 1
        ld = unk * nt if ld is None else int(ld)
        c = np.zeros((unk * ns, ld), np.float32)
        f = self.lib.o2d_oracle_panels_on_panels_coeff
        f.argtypes = [C.c_int, _fp, _fp, _up, _fp, C.c_int, _fp, _fp, _up, _fp, _fp, _fp, _fp, _fp,
                      C.c_int, C.c_int, C.c_longlong, _fp]
        rc = f(ns, _f(snx), _f(sny), _u(sidx), _f(sa), nt, _f(tnx), _f(tny), _u(tidx), _f(ta),
               _f(ttx), _f(tty), _f(tnx_), _f(tny_), int(both), int(same), ld, _f(c))
        assert rc == 0
        return c

    def vels_to_rhs_panels(self, nx, ny, idx, pu, pv, nbc=None, nunk=2):
        ttx, tty, tnx, tny, _ = self.panel_bases(nx, ny, idx)
        pu, pv = map(_f32, (pu, pv))
        nbc = np.zeros(pu.size, np.float32) if nbc is None else _f32(nbc)
        rhs = np.zeros(nunk * pu.size, np.float32)
        self.lib.o2d_oracle_vels_to_rhs_panels(int(pu.size), int(nunk), _f(pu), _f(pv), _f(ttx), _f(tty), _f(tnx),
                                               _f(tny), _f(nbc), _f(rhs))
        return rhs
    def clear_inner_layer(self, nx, ny, idx, x, y, cutoff_factor, ips, max_iter=1000):
        """clear_inner_layer, method 1 (push).  Returns new x, y and the number of pushes."""
        nx, ny = map(_f32, (nx, ny)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        _, _, snx, sny, _ = self.panel_bases(nx, ny, idx)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        f = self.lib.o2d_oracle_clear_inner_layer
        f.restype = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _up, _fp, _fp, C.c_int, _fp, _fp, C.c_float, C.c_float, C.c_int]
        moved = f(int(idx.size // 2), int(nx.size), _f(nx), _f(ny), _u(idx), _f(snx), _f(sny), int(x.size), _f(x), _f(y),
                  float(cutoff_factor), float(ips), int(max_iter))
        assert moved >= 0
        return x, y, int(moved)
    def reflect_interior(self, nx, ny, idx, x, y):
        """reflect_interior (reflect_panp2).  Returns new x, y and the number of reflected particles."""
        nx, ny = map(_f32, (nx, ny)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        _, _, snx, sny, _ = self.panel_bases(nx, ny, idx)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        f = self.lib.o2d_oracle_reflect_interior
        f.restype = C.c_longlong
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _up, _fp, _fp, C.c_int, _fp, _fp]
        moved = f(int(idx.size // 2), int(nx.size), _f(nx), _f(ny), _u(idx), _f(snx), _f(sny), int(x.size), _f(x), _f(y))
        assert moved >= 0
        return x, y, int(moved)
    def points_status(self, x, y, s):
        x, y, s = map(_f32, (x, y, s))
        out = [C.c_float() for _ in range(4)]
        self.lib.o2d_oracle_points_status(int(x.size), _f(x), _f(y), _f(s), *[C.byref(o) for o in out])
        return tuple(o.value for o in out)
    def vtu_points(self, time, x, y, s, r, u, v, w=None, e=None) -> bytes:
        """Restated Points::write_vtk: the bytes of the .vtu file."""
        x, y, u, v = map(_f32, (x, y, u, v))
        s, r, w, e = (None if a is None else _f32(a) for a in (s, r, w, e))
        self.lib.o2d_oracle_vtu_points_bound.restype = C.c_size_t
        cap = self.lib.o2d_oracle_vtu_points_bound(int(x.size))
        buf = C.create_string_buffer(cap)
        f = self.lib.o2d_oracle_vtu_points
        f.restype = C.c_longlong
        f.argtypes = [C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_double, C.c_char_p]
        nb = f(int(x.size), _f(x), _f(y), _f(s), _f(r), _f(w), _f(e), _f(u), _f(v), float(time), buf)
        assert nb > 0
        return buf.raw[:nb]
    def add_rot_strengths(self, ux, uy, idx, area, center, rotvel, vs, ss):
        ux, uy, area = map(_f32, (ux, uy, area)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        vs, ss = np.array(vs, np.float32), np.array(ss, np.float32)
        f = self.lib.o2d_oracle_add_rot_strengths
        f.argtypes = [C.c_int, _fp, _fp, _up, _fp, C.c_float, C.c_float, C.c_float, _fp, _fp]
        f(int(vs.size), _f(ux), _f(uy), _u(idx), _f(area), float(center[0]), float(center[1]), float(rotvel), _f(vs), _f(ss))
        return vs, ss

class Ref:
    """The unmodified reference templates for ONE core (oracle/_ref/libo2dref_<core>.so)."""

    kind = "reference"

    def __init__(self, core: str):
        self.core = core
        self.lib = C.CDLL(ref_path(core))
        assert self.lib.o2dref_core_id() == CORE_ID[core]

    def flops(self, which: int) -> int:
        return int(self.lib.o2dref_flops(which))

    def points_on_points(self, results, sx, sy, ss, sr, tx, ty, tr=None, tu=None, tv=None, tq=None):
        sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
        nt = tx.size
        tu = np.zeros(nt, np.float32) if tu is None else np.array(tu, np.float32)
        tv = np.zeros(nt, np.float32) if tv is None else np.array(tv, np.float32)
        want_q = results in (VELANDVORT, VELANDSHEAR)
        if want_q:
            tq = np.zeros(nt, np.float32) if tq is None else np.array(tq, np.float32)
        rc = self.lib.o2dref_points_on_points(
            int(results), int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr),
            int(nt), _f(tx), _f(ty), _f(tr), _f(tu), _f(tv), _f(tq) if want_q else None)
        if rc != 0:
            raise ValueError(f"reference rejects results type {results}")
        return (tu, tv, tq) if want_q else (tu, tv)

    def points_on_points_grad(self, sx, sy, ss, sr, tx, ty, tr, skip_self=False):
        sx, sy, ss, sr, tx, ty, tr = map(_f32, (sx, sy, ss, sr, tx, ty, tr))
        nt = tx.size
        out = [np.zeros(nt, np.float32) for _ in range(6)]
        rc = self.lib.o2dref_points_on_points_grad(
            int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr),
            int(nt), _f(tx), _f(ty), _f(tr), *[_f(o) for o in out], int(skip_self))
        assert rc == 0
        return tuple(out)

    def panels_on_points(self, nx, ny, idx, vs, ss, tx, ty, tu=None, tv=None):
        nx, ny, vs, ss, tx, ty = map(_f32, (nx, ny, vs, ss, tx, ty))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        nt = tx.size
        tu = np.zeros(nt, np.float32) if tu is None else np.array(tu, np.float32)
        tv = np.zeros(nt, np.float32) if tv is None else np.array(tv, np.float32)
        if ss is None:
            rc = self.lib.o2dref_panels_on_points_vortex(
                int(vs.size), int(nx.size), _f(nx), _f(ny), _u(idx), _f(vs),
                int(nt), _f(tx), _f(ty), _f(tu), _f(tv))
        else:
            rc = self.lib.o2dref_panels_on_points(
                int(vs.size), int(nx.size), _f(nx), _f(ny), _u(idx), _f(vs), _f(ss),
                int(nt), _f(tx), _f(ty), _f(tu), _f(tv))
        assert rc == 0
        return tu, tv

    def points_on_panels(self, sx, sy, ss, nx, ny, idx, pu=None, pv=None):
        sx, sy, ss, nx, ny = map(_f32, (sx, sy, ss, nx, ny))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        npan = idx.size // 2
        pu = np.zeros(npan, np.float32) if pu is None else np.array(pu, np.float32)
        pv = np.zeros(npan, np.float32) if pv is None else np.array(pv, np.float32)
        area = np.zeros(npan, np.float32)
        rc = self.lib.o2dref_points_on_panels(
            int(sx.size), _f(sx), _f(sy), _f(ss), int(npan), int(nx.size), _f(nx), _f(ny), _u(idx),
            _f(pu), _f(pv), _f(area))
        assert rc == 0
        return pu, pv, area

    def points_vorticity(self, sx, sy, ss, sr, tx, ty, tw=None):
        sx, sy, ss, sr, tx, ty = map(_f32, (sx, sy, ss, sr, tx, ty))
        nt = tx.size
        tw = np.zeros(nt, np.float32) if tw is None else np.array(tw, np.float32)
        rc = self.lib.o2dref_points_vorticity(
            int(sx.size), _f(sx), _f(sy), _f(ss), _f(sr), int(nt), _f(tx), _f(ty), _f(tw))
        assert rc == 0
        return tw

    def finalize_vels(self, u, v, w=None, e=None, fs=(0.0, 0.0)):
        u, v = np.array(u, np.float32), np.array(v, np.float32)
        w = None if w is None else np.array(w, np.float32)
        e = None if e is None else np.array(e, np.float32)
        self.lib.o2dref_finalize_vels.argtypes = [C.c_int, _fp, _fp, _fp, _fp, C.c_double, C.c_double]
        self.lib.o2dref_finalize_vels(int(u.size), _f(u), _f(v), _f(w), _f(e), float(fs[0]), float(fs[1]))
        return u, v, w, e

    _panels = Port._panels

    def find_vels(self, results, vort, panels, tx, ty, tr=None, fs=(0.0, 0.0)):
        tx, ty, tr = map(_f32, (tx, ty, tr))
        nt = tx.size
        vx, vy, vs, vr = map(_f32, vort) if vort is not None else (None,) * 4
        npan, nn, nx, ny, idx, pvs, pss = self._panels(panels)
        tu, tv = np.zeros(nt, np.float32), np.zeros(nt, np.float32)
        tq = np.zeros(nt, np.float32) if results in (VELANDVORT, VELANDSHEAR) else None
        f = self.lib.o2dref_find_vels
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _up, _fp, _fp,
                      C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_double, C.c_double]
        rc = f(int(results), 0 if vx is None else int(vx.size), _f(vx), _f(vy), _f(vs), _f(vr),
               npan, nn, _f(nx), _f(ny), None if idx is None else _u(idx), _f(pvs), _f(pss),
               int(nt), _f(tx), _f(ty), _f(tr), _f(tu), _f(tv), _f(tq), float(fs[0]), float(fs[1]))
        assert rc == 0
        return (tu, tv) if tq is None else (tu, tv, tq)

    def advect(self, order, x, y, s, r, dt, panels=None, fs=(0.0, 0.0)):
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r = map(_f32, (s, r))
        npan, nn, nx, ny, idx, pvs, pss = self._panels(panels)
        u, v = np.zeros(x.size, np.float32), np.zeros(x.size, np.float32)
        f = self.lib.o2dref_advect
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _up,
                      _fp, _fp, C.c_double, C.c_double, C.c_double]
        rc = f(int(order), int(x.size), _f(x), _f(y), _f(s), _f(r), _f(u), _f(v),
               npan, nn, _f(nx), _f(ny), None if idx is None else _u(idx), _f(pvs), _f(pss),
               float(fs[0]), float(fs[1]), float(dt))
        assert rc == 0
        return x, y, u, v
    def advect_bodies(self, order, x, y, s, r, panels, solve, dt, ips, fs=(0.0, 0.0), time=0.0, fldpt=None):
        """The reference's convection step with bodies (orchestration restated in ref_driver.cpp around the reference's own
        points_affect_panels / finalize_vels / find_vels pieces / Points::move / clear_inner_layer); the BEM linear solve is
        the caller's solve(stage, time, pu, pv) -> (vs, ss).  Returns (x, y, u, v, vs, ss)."""
        from omega2d_b200 import bem_solve_callback
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r = map(_f32, (s, r))
        nx, ny = map(_f32, panels[:2])
        idx = np.ascontiguousarray(panels[2], dtype=np.uint32)
        vs = np.array(panels[3], np.float32)
        ss = None if panels[4] is None else np.array(panels[4], np.float32)
        u, v = np.zeros(x.size, np.float32), np.zeros(x.size, np.float32)
        cb = bem_solve_callback(solve)
        f = self.lib.o2dref_advect_bodies
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, _fp, _up, _fp, _fp,
                      C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, C.c_int, _fp, _fp, _fp, _fp, C.c_void_p, C.c_void_p]
        nf = 0
        fx = fy = fu = fv = None
        if fldpt is not None:
            fx, fy = np.array(fldpt[0], np.float32), np.array(fldpt[1], np.float32)
            nf = fx.size
            fu, fv = np.zeros(nf, np.float32), np.zeros(nf, np.float32)
        rc = f(int(order), int(x.size), _f(x), _f(y), _f(s), _f(r), _f(u), _f(v), int(vs.size), int(nx.size), _f(nx), _f(ny),
               _u(idx), _f(vs), _f(ss), float(fs[0]), float(fs[1]), float(time), float(dt), float(ips),
               nf, _f(fx), _f(fy), _f(fu), _f(fv), C.cast(cb, C.c_void_p), None)
        assert rc == 0, rc
        return (x, y, u, v, vs, ss) if fldpt is None else (x, y, u, v, vs, ss, fx, fy, fu, fv)

    def panels_on_panels_coeff(self, src, targ=None):
        """The reference's own panels_on_panels_coeff<float> (always both strengths).  Returns
        (coeffs as (2*ns, 2*nt) array, (ttx, tty, tnx, tny, tarea), sarea)."""
        same = targ is None
        targ = src if same else targ
        snx, sny = map(_f32, src[:2]); sidx = np.ascontiguousarray(src[2], dtype=np.uint32)
        tnx, tny = map(_f32, targ[:2]); tidx = np.ascontiguousarray(targ[2], dtype=np.uint32)
        ns, nt = sidx.size // 2, tidx.size // 2
        c = np.zeros((2 * ns, 2 * nt), np.float32)
        bases = [np.zeros(nt, np.float32) for _ in range(5)]
        sa = np.zeros(ns, np.float32)
        rc = self.lib.o2dref_panels_on_panels_coeff(ns, int(snx.size), _f(snx), _f(sny), _u(sidx),
                                                    nt, int(tnx.size), _f(tnx), _f(tny), _u(tidx), int(same),
                                                    _f(c), *[_f(b) for b in bases], _f(sa))
        assert rc == 0
        return c, tuple(bases), sa

    def panels_on_panels_coeff_vortex(self, src, targ, bases, sa, same):
        snx, sny = map(_f32, src[:2]); sidx = np.ascontiguousarray(src[2], dtype=np.uint32)
        tnx, tny = map(_f32, targ[:2]); tidx = np.ascontiguousarray(targ[2], dtype=np.uint32)
        ns, nt = sidx.size // 2, tidx.size // 2
        c = np.zeros((ns, nt), np.float32)
        rc = self.lib.o2dref_panels_on_panels_coeff_vortex(ns, _f(snx), _f(sny), _u(sidx), _f(_f32(sa)),
                                                           nt, _f(tnx), _f(tny), _u(tidx), _f(_f32(bases[4])),
                                                           _f(_f32(bases[0])), _f(_f32(bases[1])), int(same), _f(c))
        assert rc == 0
        return c

    def vels_to_rhs_panels(self, nx, ny, idx, pu, pv, nbc=None):
        nx, ny, pu, pv, nbc = map(_f32, (nx, ny, pu, pv, nbc))
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        rhs = np.zeros(2 * pu.size, np.float32)
        n = self.lib.o2dref_vels_to_rhs_panels(int(pu.size), int(nx.size), _f(nx), _f(ny), _u(idx), _f(pu), _f(pv),
                                               _f(nbc), _f(rhs))
        assert n == rhs.size
        return rhs
    def clear_inner_layer(self, nx, ny, idx, x, y, cutoff_factor, ips, s=None, r=None, method=1):
        nx, ny = map(_f32, (nx, ny)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s = None if s is None else np.array(s, np.float32)
        r = None if r is None else _f32(r)
        f = self.lib.o2dref_clear_inner_layer
        f.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _up, C.c_int, _fp, _fp, _fp, _fp, C.c_float, C.c_float]
        rc = f(int(method), int(idx.size // 2), int(nx.size), _f(nx), _f(ny), _u(idx), int(x.size), _f(x), _f(y), _f(s), _f(r),
               float(cutoff_factor), float(ips))
        assert rc == 0
        return (x, y) if s is None else (x, y, s)
    def reflect_interior(self, nx, ny, idx, x, y, s, r):
        nx, ny = map(_f32, (nx, ny)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        x, y = np.array(x, np.float32), np.array(y, np.float32)
        s, r = _f32(s), _f32(r)
        f = self.lib.o2dref_reflect_interior
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _up, C.c_int, _fp, _fp, _fp, _fp]
        rc = f(int(idx.size // 2), int(nx.size), _f(nx), _f(ny), _u(idx), int(x.size), _f(x), _f(y), _f(s), _f(r))
        assert rc == 0
        return x, y
    def write_vtk_points(self, directory, index, frameno, time, x, y, s, r, u, v, w=None, e=None):
        """The reference's own Points::write_vtk.  Returns the bytes of the .vtu file it wrote and its name."""
        x, y, u, v = map(_f32, (x, y, u, v))
        s = None if s is None else _f32(s); r = None if r is None else _f32(r); w = None if w is None else _f32(w)
        e = None if e is None else _f32(e)
        f = self.lib.o2dref_write_vtk_points
        f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_char_p, C.c_int]
        buf = C.create_string_buffer(256)
        rc = f(str(directory).encode(), int(index), int(frameno), float(time), int(x.size), _f(x), _f(y), _f(s), _f(r),
               _f(u), _f(v), _f(w), _f(e), buf, 256)
        assert rc == 0, rc
        name = buf.value.decode()
        with open(os.path.join(str(directory), name), "rb") as fh:
            return fh.read(), name
    def add_rot_strengths(self, nx, ny, idx, rotvel, vs, ss):
        """-> (vs, ss, (cx, cy)) from the reference's own body-bound Surfaces."""
        nx, ny = map(_f32, (nx, ny)); idx = np.ascontiguousarray(idx, dtype=np.uint32)
        vs, ss = np.array(vs, np.float32), np.array(ss, np.float32)
        cx, cy = C.c_float(), C.c_float()
        f = self.lib.o2dref_add_rot_strengths
        f.argtypes = [C.c_int, C.c_int, _fp, _fp, _up, C.c_float, _fp, _fp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        rc = f(int(vs.size), int(nx.size), _f(nx), _f(ny), _u(idx), float(rotvel), _f(vs), _f(ss), C.byref(cx), C.byref(cy))
        assert rc == 0
        return vs, ss, (cx.value, cy.value)

def host_cores() -> int:
    """Cores this process may run on (what OpenMP would use by default)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n: int | None = None) -> int:
    """Set the OpenMP team size of the checkers.  torch.distributed.run exports
    OMP_NUM_THREADS=1 into every rank; the CPU baseline must still use all host cores."""
    n = n or host_cores()
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
    return n


_port = None
_refs = {}


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref(core: str) -> Ref:
    if core not in _refs:
        _refs[core] = Ref(core)
    return _refs[core]
