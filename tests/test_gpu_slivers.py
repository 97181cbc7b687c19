"""GPU parity on meshes made of slivers: the Delaunay triangulation of point CLUSTERS (every site is 3-5 points a few 1e-7 .. 1e-6
apart), so that the long thin tetrahedra between two clusters have nearly coincident faces and a ray crosses several of them within
eps = 1e-6 in t.  The literal dedupe / pairing of post_process_tetrahedra (optix_trace_rays.cu:110-266) then does real work on
every ray (marks, deletions, swaps), which random point clouds only touch on 5-6 % of the rays.  Bit-exact against the oracle on
every implementation of trace_rays; this is the stress test of the windowed pairing (tn_trace.cu: post_process_windows), swap
cascades included.  tests/test_window_pairing_model.py checks the windowing argument itself on the CPU."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn

KEYS = ["num_visited_cells", "visited_cells", "vertex_indices", "hit_distances", "barycentric_coordinates"]


def sliver_mesh(sites=500, per=3, jitter=2e-6, seed=5):
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    S = rng.random((sites, 3))
    P = (S[:, None, :] + jitter * rng.standard_normal((sites, per, 3))).reshape(-1, 3)
    V = np.unique(P.astype(np.float32), axis=0)
    C = Delaunay(V.astype(np.float64)).simplices.astype(np.int32)
    return V, C


def test_sliver_mesh_oracle_is_sane():
    """CPU: the fixture produces rays with many sub-eps crossings (otherwise the GPU test below tests nothing)"""
    V, C = sliver_mesh()
    o, d = syn.camera_rays(64, seed=2)
    ref = orc.OracleMesh(V, C).trace_rays(o, d, 512)
    n = ref["num_visited_cells"]
    assert (n > 10).sum() > 32
    t = ref["hit_distances"]
    short = sum(int((np.diff(t[r, : n[r], 0]) < 2e-6).sum()) for r in range(len(o)))
    assert short > 150, short


@pytest.mark.gpu
@pytest.mark.parametrize("jitter,per,gen", [(2e-6, 3, "camera"), (5e-7, 4, "sphere"), (1e-6, 5, "camera"), (4e-6, 3, "sphere")])
def test_sliver_mesh_bit_exact(jitter, per, gen):
    from conftest import TRACE_IMPLS, force_trace_impl
    from tetranerf import cpp
    from tetranerf.utils.extension import tetranerf_cpp_extension as ext

    dev = torch.device("cuda:0")
    V, C = sliver_mesh(jitter=jitter, per=per)
    o, d = (syn.camera_rays if gen == "camera" else syn.sphere_rays)(1500, seed=7)
    ref = orc.OracleMesh(V, C).trace_rays(o, d, 512)
    tr = cpp.TetrahedraTracer(dev)
    tr.load_tetrahedra(torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev))
    for impl in TRACE_IMPLS:
        force_trace_impl(tr, impl)
        out = tr.trace_rays(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), 512)
        tr.synchronize()
        walkable, listed = tr.trace_stats()
        print(f"jitter {jitter} per {per} {impl}: walkable {walkable}, listed {listed}, all-hits {ext._lib.tn_debug_last_exact_count()}, "
              f"records/ray {ref['num_visited_cells'].mean():.1f}")
        for k in KEYS:
            a, b = out[k].cpu().numpy(), ref[k]
            if a.dtype.kind == "f":
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), f"{impl}: {k} differs on rays {np.unique(np.nonzero(a != b)[0])[:10]}"
