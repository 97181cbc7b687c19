"""CPU: the C-ABI library loads and exports every symbol include/tetranerf_b200.h declares; without a CUDA
device the product fails loudly (no fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "tetra-nerf_b200" / "csrc" / "libtetranerf_b200.so"


def _declared():
    text = (ROOT / "include" / "tetranerf_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert LIB.exists(), "run python tetra-nerf_b200/build.py"
    lib = ctypes.CDLL(str(LIB))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tetranerf_b200.h but not exported"


def test_shim_surface_matches_reference_pybind():
    """src/py_binding.cpp:433-449"""
    from tetranerf import cpp
    from tetranerf.utils import extension

    for n in ("trace_rays", "trace_rays_triangles", "find_visited_cells", "find_tetrahedra", "load_tetrahedra", "device"):
        assert hasattr(cpp.TetrahedraTracer, n)
    for n in ("triangulate", "find_average_spacing", "interpolate_values", "interpolate_values_backward", "gather_uint32", "scatter_ema_uint32"):
        assert callable(getattr(cpp, n))
    for n in ("TetrahedraTracer", "triangulate", "gather_uint32", "scatter_ema_uint32_", "interpolate_values", "add_barycentrics_grad"):
        assert hasattr(extension, n)


def test_non_cuda_device_raises():
    from tetranerf import cpp

    with pytest.raises(RuntimeError, match="CUDA device"):  # py_binding.cpp:31-33
        cpp.TetrahedraTracer(torch.device("cpu"))


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without GPU")
def test_no_silent_cpu_fallback():
    from tetranerf import cpp

    with pytest.raises(RuntimeError):
        cpp.TetrahedraTracer(torch.device("cuda:0"))
    with pytest.raises(RuntimeError):
        cpp.interpolate_values(torch.zeros((1, 4), dtype=torch.int32), torch.zeros((1, 3)), torch.zeros((2, 3)))


def test_pybind_module_surface():
    """the compiled pybind11 module named like the reference's (src/py_binding.cpp:433-449) imports without a GPU and exposes the surface"""
    import importlib.util
    import sys as _sys

    bp = ROOT / "tetra-nerf_b200" / "build.py"
    spec = importlib.util.spec_from_file_location("tn_build_for_test", bp)
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.pybind_path().exists(), "run python tetra-nerf_b200/build.py"
    from tetranerf.utils.extension._pybind import tetranerf_cpp_extension as pb

    assert pb.__name__.endswith("tetranerf_cpp_extension") and pb.BINDING == "pybind11"
    for n in ("trace_rays", "trace_rays_triangles", "find_visited_cells", "find_tetrahedra", "load_tetrahedra", "device"):
        assert hasattr(pb.TetrahedraTracer, n)
    for n in ("triangulate", "find_average_spacing", "interpolate_values", "interpolate_values_backward", "gather_uint32", "scatter_ema_uint32"):
        assert callable(getattr(pb, n))
    with pytest.raises(RuntimeError, match="CUDA device"):
        pb.TetrahedraTracer(torch.device("cpu"))
