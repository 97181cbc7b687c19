"""GPU: the compiled pybind11 module `tetranerf_cpp_extension` (csrc/py_binding.cpp; reference src/py_binding.cpp:433-449) against the
ctypes shim of the same name -- same surface, bit-identical tensors, same RuntimeErrors -- and through the fused renderer."""
import numpy as np
import pytest
import torch

from tetranerf.b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _both():
    from tetranerf.utils.extension import tetranerf_cpp_extension as ct
    from tetranerf.utils.extension._pybind import tetranerf_cpp_extension as pb

    assert pb.BINDING == "pybind11"
    return ct, pb


def test_pybind_matches_ctypes_bitwise(small_mesh):
    ct, pb = _both()
    V, C = small_mesh
    dV, dC = torch.from_numpy(V).to(DEV), torch.from_numpy(C).to(DEV)
    a, b = ct.TetrahedraTracer(DEV), pb.TetrahedraTracer(DEV)
    assert b.device == DEV
    a.load_tetrahedra(dV, dC); b.load_tetrahedra(dV, dC)
    assert a.num_faces() == b.num_faces()
    o, d = syn.camera_rays(700, seed=4)
    o[3] = [5, 5, 5]
    do, dd = torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV)
    ra, rb = a.trace_rays(do, dd, 256), b.trace_rays(do, dd, 256)
    a.synchronize(); b.synchronize()
    assert set(ra) == set(rb)
    for k in ra:
        assert ra[k].dtype == rb[k].dtype and ra[k].shape == rb[k].shape and torch.equal(ra[k].view(torch.int32), rb[k].view(torch.int32)), k
    ta, tb = a.trace_rays_triangles(do, dd, 256), b.trace_rays_triangles(do, dd, 256)
    for k in ta:
        assert torch.equal(ta[k].view(torch.int32), tb[k].view(torch.int32)), k
    pos = torch.rand((4, 50, 3), device=DEV)
    fa, fb = a.find_tetrahedra(pos), b.find_tetrahedra(pos)
    for k in fa:
        assert fa[k].shape == fb[k].shape and torch.equal(fa[k], fb[k]), k
    keep = ra["num_visited_cells"] > 0
    dist = (torch.rand((int(keep.sum()), 40), device=DEV).sort(-1).values * 3.0).contiguous()
    args = [ra[k][keep].contiguous() for k in ("num_visited_cells", "visited_cells", "barycentric_coordinates", "hit_distances", "vertex_indices")]
    ma, mb = a.find_visited_cells(*args, dist), b.find_visited_cells(*args, dist)
    for k in ma:
        assert ma[k].dtype == mb[k].dtype and torch.equal(ma[k], mb[k]), k
    field = torch.randn((64, len(V)), device=DEV)
    ia = ct.interpolate_values(ma["vertex_indices"], ma["barycentric_coordinates"], field)
    ib = pb.interpolate_values(ma["vertex_indices"], ma["barycentric_coordinates"], field)
    assert ia.shape == ib.shape and torch.equal(ia, ib)
    g = torch.randn_like(ia)
    ga = ct.interpolate_values_backward(ma["vertex_indices"], ma["barycentric_coordinates"], field, g)
    gb = pb.interpolate_values_backward(ma["vertex_indices"], ma["barycentric_coordinates"], field, g)
    assert ga.shape == gb.shape == (64, len(V))
    assert (ga - gb).abs().max().item() <= 1e-5 * ga.abs().max().item()  # atomics: order differs run to run


def test_pybind_errors_are_runtime_errors(cube_mesh):
    _, pb = _both()
    with pytest.raises(RuntimeError, match="CUDA device"):  # py_binding.cpp:31-33
        pb.TetrahedraTracer(torch.device("cpu"))
    V, C = cube_mesh
    t = pb.TetrahedraTracer(DEV)
    t.load_tetrahedra(torch.from_numpy(V).to(DEV), torch.from_numpy(C).to(DEV))
    o = torch.zeros((2, 3), device=DEV)
    with pytest.raises(RuntimeError, match="power of 2"):  # py_binding.cpp:44-47
        t.trace_rays(o, o, 100)
    with pytest.raises(RuntimeError, match="float32"):
        t.trace_rays(o.double(), o, 16)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        t.trace_rays(o.cpu(), o, 16)
    with pytest.raises(RuntimeError, match="Unsupported interpolation dimension"):
        pb.interpolate_values(torch.zeros((1, 5), dtype=torch.int32, device=DEV), torch.zeros((1, 4), device=DEV), torch.zeros((2, 3), device=DEV))
    with pytest.raises(RuntimeError):
        pb.gather_uint32(torch.zeros(1), 0, torch.zeros(1, dtype=torch.int32))


def test_fused_renderer_over_the_pybind_tracer(small_mesh):
    ct, pb = _both()
    from oracle import oracle as orc
    from tetranerf.b200.render import FusedRenderer, RenderSettings

    V, C = small_mesh
    field = syn.random_field(len(V), 64, seed=3)
    o, d = syn.camera_rays(128, seed=6)
    outs = []
    for mod in (ct, pb):
        tr = mod.TetrahedraTracer(DEV)
        tr.load_tetrahedra(torch.from_numpy(V).to(DEV), torch.from_numpy(C).to(DEV))
        fr = FusedRenderer(tr)
        fr.set_field(torch.from_numpy(field).to(DEV))
        fr.set_weights(orc.init_mlp_params(0))
        out = fr.render(torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), RenderSettings.tetra_nerf())
        tr.synchronize()
        outs.append({k: v.clone() for k, v in out.items()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
