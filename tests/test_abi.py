"""CPU: the C-ABI library loads and exports every symbol include/omega2d_b200.h declares; with no
GPU it fails loudly (no CPU fallback); its flop accounting equals the reference's counters."""
import ctypes
import re

import numpy as np

import pytest

import omega2d_b200 as o2d
import oracle


def declared_functions():
    text = open(o2d.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(o2d_[a-z0-9_]+|external_vel_solver_[a-z_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported():
    lib = o2d.load_library()
    names = declared_functions()
    assert len(names) >= 24, names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/omega2d_b200.h but not exported"


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(o2d.O2DError) as ei:
        o2d.Context(0)
    assert ei.value.code == o2d.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
    # the legacy plugin symbols report failure through a negative flop count
    import numpy as np
    a = np.zeros(4, np.float32)
    with pytest.raises(o2d.O2DError):
        o2d.external_vel_solver(a, a, a, a + 1, a, a)


@pytest.mark.parametrize("core", oracle.CORES)
def test_flops_match_reference_counters(core, port):
    # oracle 'which': 0 u, 1 uw, 2 ue, 3 ug ; library results_t: 1, 4, 5, 2
    for which, res in ((0, o2d.VELONLY), (1, o2d.VELANDVORT), (2, o2d.VELANDSHEAR)):
        for blob in (False, True):
            assert o2d.flops_per_pair(0, core, res, blob) == port.flops(core, which, blob)
    assert o2d.flops_per_pair(0, core, o2d.VELANDGRAD, True) == port.flops(core, 3, True)
    assert o2d.flops_per_pair(1) == port.flops(core, 4, False) == 39
    assert o2d.flops_per_pair(2) == port.flops(core, 5, False) == 47
    assert o2d.flops_per_pair(3) == port.flops(core, 6, False) == 13


def test_default_core_is_the_references_default():
    lib = o2d.load_library()
    assert lib.o2d_get_default_core() == o2d.CORE_ID["v2"]   # reference src/CoreFunc.h:38
    assert lib.o2d_set_default_core(99) == o2d.ERR_INVALID


def test_group_without_gpu_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(o2d.O2DError) as ei:
        o2d.Group()
    assert ei.value.code == o2d.ERR_NO_DEVICE


@pytest.mark.parametrize("n", [1, 2, 3, 11, 12, 13, 1000, 12345])
def test_vtu_size_query_matches_the_restated_writer(n, port):
    """o2d_vtu_points_size is pure host logic (the XML skeleton + base64 lengths): it must equal the length of the file
    the restated Points::write_vtk produces, for every combination of optional arrays."""
    import itertools
    a = np.zeros(n, np.float32)
    for hs, hr, hw, he in itertools.product((0, 1), repeat=4):
        img = port.vtu_points(0.5, a, a, a if hs else None, a if hr else None, a, a, w=a if hw else None, e=a if he else None)
        assert len(img) == o2d.vtu_points_size(n, hs, hr, hw, he), (n, hs, hr, hw, he)
    with pytest.raises(o2d.O2DError):
        o2d.vtu_points_size(0)


def test_cmake_build_exports_the_same_abi(tmp_path):
    """omega2d_b200/CMakeLists.txt (what a CMake host project such as Omega2D adds with add_subdirectory) builds the same
    library: every function the header declares is exported.  nvcc cross-compiles sm_100a without a GPU."""
    import os
    import shutil
    import subprocess
    if not (shutil.which("cmake") and shutil.which("ninja") and os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("cmake / ninja / nvcc not available")
    src = os.path.join(os.path.dirname(o2d.HEADER_PATH), "..", "omega2d_b200")
    cfg = subprocess.run(["cmake", "-G", "Ninja", "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc", "-DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++",
                          "-DCMAKE_CXX_COMPILER=/usr/bin/g++", "-S", src, "-B", str(tmp_path)], capture_output=True, text=True)
    assert cfg.returncode == 0, cfg.stdout[-2000:] + cfg.stderr[-2000:]
    bld = subprocess.run(["ninja", "-C", str(tmp_path)], capture_output=True, text=True)
    assert bld.returncode == 0, bld.stdout[-3000:]
    lib = ctypes.CDLL(str(tmp_path / "libomega2d_b200.so"))
    for n in declared_functions():
        assert hasattr(lib, n), n
