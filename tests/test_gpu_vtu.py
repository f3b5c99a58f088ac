"""GPU: o2d_vtu_points / o2d_write_vtu_points against the reference-written golden .vtu files and the restated writer,
byte for byte; size-independent decode-back check at 1M particles."""
import numpy as np
import pytest

import omega2d_b200 as o2d
from test_vtu_oracle import decode_arrays, golden_case, gv  # noqa: F401  (gv is a fixture)

pytestmark = pytest.mark.gpu


def test_golden_files(ctx, gv):
    for tag in gv["tags"]:
        kw, want = golden_case(gv, str(tag))
        n = kw["x"].size
        got = ctx.vtu_points(o2d.MEM_HOST, n, kw["x"], kw["y"], kw["s"], kw["r"], kw["u"], kw["v"], kw["time"], w=kw["w"], e=kw["e"])
        assert got == want, tag


@pytest.mark.parametrize("n", list(range(1, 40)) + [255, 256, 257, 3071, 3072, 3073, 100003])
def test_every_tail_vs_oracle(n, ctx, port):
    rng = np.random.default_rng(n)
    a = [rng.standard_normal(n).astype(np.float32) for _ in range(8)]
    a[0].view(np.uint32)[:] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)   # arbitrary bit patterns (NaNs too)
    for s, r, w, e in ((a[2], a[3], None, None), (None, None, None, None), (a[2], a[3], a[6], a[7])):
        want = port.vtu_points(0.75, a[0], a[1], s, r, a[4], a[5], w=w, e=e)
        got = ctx.vtu_points(o2d.MEM_HOST, n, a[0], a[1], s, r, a[4], a[5], 0.75, w=w, e=e)
        assert got == want


def test_device_arrays_and_file(ctx, port, tmp_path):
    import torch
    n = 54321
    rng = np.random.default_rng(1)
    a = [rng.standard_normal(n).astype(np.float32) for _ in range(7)]
    d = [torch.from_numpy(x).cuda() for x in a]
    want = port.vtu_points(2.0, a[0], a[1], a[2], a[3], a[4], a[5], w=a[6])
    got = ctx.vtu_points(o2d.MEM_DEVICE, n, d[0], d[1], d[2], d[3], d[4], d[5], 2.0, w=d[6])
    assert got == want
    path = tmp_path / "part_00_00001.vtu"
    ctx.write_vtu_points(o2d.MEM_DEVICE, path, n, d[0], d[1], d[2], d[3], d[4], d[5], 2.0, w=d[6])
    assert path.read_bytes() == want
    ctx.write_vtu_points(o2d.MEM_HOST, path, 5, a[0][:5].copy(), a[1][:5].copy(), None, None, a[4][:5].copy(), a[5][:5].copy(), 0.0)
    assert path.read_bytes() == port.vtu_points(0.0, a[0][:5], a[1][:5], None, None, a[4][:5], a[5][:5])


def test_one_million_decodes_back(ctx):
    n = 1_000_000
    rng = np.random.default_rng(9)
    x, y, s, r, u, v = (rng.standard_normal(n).astype(np.float32) for _ in range(6))
    img = ctx.vtu_points(o2d.MEM_HOST, n, x, y, s, r, u, v, 0.5)
    assert len(img) == o2d.vtu_points_size(n) == 54668004
    arr = decode_arrays(img)
    pos = np.frombuffer(arr["position"], "<f4").reshape(n, 3)
    vel = np.frombuffer(arr["velocity"], "<f4").reshape(n, 3)
    assert (pos[:, 0] == x).all() and (pos[:, 1] == y).all() and (pos[:, 2] == 0).all()
    assert (vel[:, 0] == u).all() and (vel[:, 1] == v).all() and (vel[:, 2] == 0).all()
    assert (np.frombuffer(arr["connectivity"], "<i4") == np.arange(n)).all()
    assert (np.frombuffer(arr["offsets"], "<i4") == np.arange(1, n + 1)).all()
    assert arr["types"] == b"\x01" * n
    assert (np.frombuffer(arr["circulation"], "<f4") == s).all() and (np.frombuffer(arr["radius"], "<f4") == r).all()


def test_errors(ctx):
    a = np.zeros(4, np.float32)
    with pytest.raises(o2d.O2DError) as ei:
        ctx.vtu_points(o2d.MEM_HOST, 4, a, a, a, a, a, a, 0.0, cap=100)
    assert ei.value.code == o2d.ERR_INVALID
    with pytest.raises(o2d.O2DError):
        o2d.vtu_points_size(0)
    with pytest.raises(o2d.O2DError) as ei:
        o2d.vtu_points_size(1 << 28)
    assert ei.value.code == o2d.ERR_UNSUPPORTED
    with pytest.raises(o2d.O2DError):
        ctx.write_vtu_points(o2d.MEM_HOST, "/nonexistent-dir/x.vtu", 4, a, a, a, a, a, a, 0.0)
