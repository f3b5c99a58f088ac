"""TEST INFRASTRUCTURE (like everything under oracle/): the reference's panel formulas (kernelu_1vs_0p /
kernelu_1v_0p, reference src/Kernels.h:312-405, get_ustar :243-269) evaluated in numpy float64.

Why it exists: the reference evaluates `atan2f - atan2f` and `logf` of a ratio in float32, and for short panels
(a finely panelled body seen from a few diameters away) that form loses more than the 1e-5 parity tolerance to
cancellation -- the reference's own result is 2e-5 ... 4e-4 away from exact arithmetic.  Where that happens a
float32 result cannot be "within 1e-5 of the reference" and correct at the same time, so the parity rule for the
panel kernels is stated against BOTH (README "Parity rule for panels"):
    err(gpu, reference) <= 1e-5,  or -- where err(reference, float64) > 1e-5 --
    err(gpu, float64) <= 1e-5  and  err(gpu, reference) <= 1.5 * err(reference, float64).
These functions supply the float64 side; fixtures made with them are committed under tests/golden/."""
import numpy as np


def panels_on_points_f64(nx, ny, idx, vs, ss, tx, ty):
    nx, ny, tx, ty = (np.asarray(a, np.float64) for a in (nx, ny, tx, ty))
    vs = np.asarray(vs, np.float64)
    ss = np.zeros_like(vs) if ss is None else np.asarray(ss, np.float64)
    i0, i1 = idx[0::2].astype(np.int64), idx[1::2].astype(np.int64)
    u = np.zeros(tx.size); v = np.zeros(tx.size)
    for lo in range(0, tx.size, 512):
        X = tx[lo:lo + 512, None]; Y = ty[lo:lo + 512, None]
        dx0 = X - nx[i0][None, :]; dy0 = Y - ny[i0][None, :]
        dx1 = X - nx[i1][None, :]; dy1 = Y - ny[i1][None, :]
        px = (nx[i1] - nx[i0])[None, :]; py = (ny[i1] - ny[i0])[None, :]
        ln = np.sqrt(px * px + py * py); px = px / ln; py = py / ln
        ustar = np.arctan2(dx1, dy1) - np.arctan2(dx0, dy0)
        ustar = np.where(ustar < -np.pi, ustar + 2 * np.pi, ustar)
        ustar = np.where(ustar > np.pi, ustar - 2 * np.pi, ustar)
        vstar = 0.5 * np.log((dx0 * dx0 + dy0 * dy0) / (dx1 * dx1 + dy1 * dy1))
        u[lo:lo + 512] = (vs * (ustar * px - vstar * py) + ss * (ustar * py + vstar * px)).sum(axis=1)
        v[lo:lo + 512] = (vs * (ustar * py + vstar * px) + ss * (vstar * py - ustar * px)).sum(axis=1)
    return u, v


def points_on_panels_f64(sx, sy, ss, nx, ny, idx, area):
    """Exact-arithmetic points_affect_panels (reference src/Influence.h:719-763)."""
    sx, sy, ss, nx, ny, area = (np.asarray(a, np.float64) for a in (sx, sy, ss, nx, ny, area))
    i0, i1 = idx[0::2].astype(np.int64), idx[1::2].astype(np.int64)
    npan = i0.size
    pu = np.zeros(npan); pv = np.zeros(npan)
    for lo in range(0, npan, 256):
        a0, a1 = i0[lo:lo + 256], i1[lo:lo + 256]
        X = sx[None, :]; Y = sy[None, :]
        dx0 = X - nx[a0][:, None]; dy0 = Y - ny[a0][:, None]
        dx1 = X - nx[a1][:, None]; dy1 = Y - ny[a1][:, None]
        px = (nx[a1] - nx[a0])[:, None]; py = (ny[a1] - ny[a0])[:, None]
        ln = np.sqrt(px * px + py * py)
        ustar = np.arctan2(dx1, dy1) - np.arctan2(dx0, dy0)
        ustar = np.where(ustar < -np.pi, ustar + 2 * np.pi, ustar)
        ustar = np.where(ustar > np.pi, ustar - 2 * np.pi, ustar)
        vstar = 0.5 * np.log((dx0 * dx0 + dy0 * dy0) / (dx1 * dx1 + dy1 * dy1))
        mult = ss[None, :] / ln
        pu[lo:lo + 256] = -(mult * (ustar * px - vstar * py)).sum(axis=1) / area[lo:lo + 256]
        pv[lo:lo + 256] = -(mult * (ustar * py + vstar * px)).sum(axis=1) / area[lo:lo + 256]
    return pu, pv
