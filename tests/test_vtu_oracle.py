"""CPU: the restated Points::write_vtk (oracle/o2d_oracle.c) against the reference-written golden .vtu files and,
where oracle/_ref exists, against the reference's own writer on fresh inputs; plus format properties that hold at
any size (every DataArray decodes back to its input; the header is the length of the encoded text)."""
import base64
import os
import re
import tempfile

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gv():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_vtu.npz"))


def golden_case(gv, tag):
    x, y, s, r, u, v, w, e = gv[f"{tag}_in"]
    blobs, hw, he = gv[f"{tag}_flags"]
    return dict(time=float(gv[f"{tag}_time"]), x=x, y=y, s=s if blobs else None, r=r if blobs else None, u=u, v=v,
                w=w if hw else None, e=e if he else None), gv[f"{tag}_file"].tobytes()


def decode_arrays(img: bytes):
    """-> {Name: raw bytes} of every DataArray; checks the UInt32 header quirk on the way."""
    out = {}
    for m in re.finditer(rb'<DataArray Name="(\w+)"[^>]*format="binary"> (\S+) </DataArray>', img):
        text = m.group(2)
        enc_len = int(np.frombuffer(base64.b64decode(text[:8]), "<u4")[0])
        assert enc_len == len(text) - 8          # length of the ENCODED text (src/VtkXmlWriter.h:99-100)
        out[m.group(1).decode()] = base64.b64decode(text[8:])
    return out


def test_port_matches_golden_files(gv, port):
    for tag in gv["tags"]:
        kw, want = golden_case(gv, str(tag))
        assert port.vtu_points(**kw) == want, tag


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [1, 6, 35, 36, 37, 4099])
def test_port_matches_reference_writer(n, port):
    R = oracle.ref("v2")
    rng = np.random.default_rng(n)
    a = [rng.standard_normal(n).astype(np.float32) for _ in range(8)]
    with tempfile.TemporaryDirectory() as d:
        for s, r, w, e in ((a[2], a[3], None, None), (None, None, None, None), (a[2], a[3], a[6], a[7]), (None, None, a[6], None)):
            want, _ = R.write_vtk_points(d, 0, 1, 0.25, a[0], a[1], s, r, a[4], a[5], w=w, e=e)
            assert port.vtu_points(0.25, a[0], a[1], s, r, a[4], a[5], w=w, e=e) == want


def test_arrays_decode_back(port):
    n = 10007
    rng = np.random.default_rng(5)
    x, y, s, r, u, v, w = (rng.standard_normal(n).astype(np.float32) for _ in range(7))
    arr = decode_arrays(port.vtu_points(1.5, x, y, s, r, u, v, w=w))
    pos = np.frombuffer(arr["position"], "<f4").reshape(n, 3)
    assert (pos[:, 0] == x).all() and (pos[:, 1] == y).all() and (pos[:, 2] == 0).all()
    vel = np.frombuffer(arr["velocity"], "<f4").reshape(n, 3)
    assert (vel[:, 0] == u).all() and (vel[:, 1] == v).all() and (vel[:, 2] == 0).all()
    assert (np.frombuffer(arr["connectivity"], "<i4") == np.arange(n)).all()
    assert (np.frombuffer(arr["offsets"], "<i4") == np.arange(1, n + 1)).all()
    assert arr["types"] == b"\x01" * n
    for name, a in (("circulation", s), ("radius", r), ("vorticity", w)):
        assert (np.frombuffer(arr[name], "<f4") == a).all()
    assert np.frombuffer(arr["TimeValue"], "<f8")[0] == 1.5
