"""Deterministic synthetic inputs for the ray-sampling hot path (no datasets are available offline).

Meshes follow SURVEY.md §8d: uniform random points in the unit cube -> scipy Delaunay (the reference
uses CGAL, src/triangulation.cpp:34-75; cell order/orientation is irrelevant once the mesh is an
input).  Rays: camera-like coherent bundle, or incoherent (origins on a radius-2 sphere)."""
from __future__ import annotations

import numpy as np

CUBE_VERTICES = np.array(
    [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1], [0.5, 0.5, 0.5]], dtype=np.float32
)
# the 12-tetrahedra cube of the reference's tests/test_tetrahedra_tracer.py:231-253
CUBE_CELLS = np.array(
    [[0, 1, 2, 8], [2, 1, 3, 8], [0, 1, 4, 8], [4, 1, 5, 8], [0, 2, 4, 8], [4, 2, 6, 8],
     [4, 5, 6, 8], [5, 6, 7, 8], [2, 3, 6, 8], [3, 6, 7, 8], [1, 3, 5, 8], [3, 5, 7, 8]], dtype=np.int32
)


def delaunay_mesh(num_points: int, seed: int = 0):
    """-> (vertices f32[V,3], cells i32[T,4]).  45_000 points -> 302,024 tets; 150_000 -> 1,009,158."""
    from scipy.spatial import Delaunay

    pts = np.random.default_rng(seed).random((num_points, 3), dtype=np.float32)
    cells = Delaunay(pts.astype(np.float64)).simplices.astype(np.int32)
    return pts, np.ascontiguousarray(cells)


def camera_rays(num_rays: int, seed: int = 1):
    """Coherent bundle: origins (0.5,-1.5,0.5)+N(0,0.05^2), targets uniform in [0.2,0.8]^3, unit directions."""
    rng = np.random.default_rng(seed)
    o = (np.array([0.5, -1.5, 0.5]) + 0.05 * rng.standard_normal((num_rays, 3))).astype(np.float32)
    tgt = (0.2 + 0.6 * rng.random((num_rays, 3))).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


def sphere_rays(num_rays: int, seed: int = 2, radius: float = 2.0):
    """Incoherent: origins uniform on a sphere of given radius around the cube centre, random targets inside."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((num_rays, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    o = (0.5 + radius * v).astype(np.float32)
    tgt = (0.1 + 0.8 * rng.random((num_rays, 3))).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return o, d


def random_field(num_vertices: int, field_dim: int = 64, seed: int = 3, kind: str = "normal"):
    """tetrahedra_field [field_dim, V] (feature-major, model.py:247-255).  kind "init" = U(-1e-4,1e-4)
    as model.py:268-271; "normal" = N(0,1) so that parity is not vacuous."""
    rng = np.random.default_rng(seed)
    if kind == "init":
        return ((rng.random((field_dim, num_vertices)) * 2 - 1) * 1e-4).astype(np.float32)
    return rng.standard_normal((field_dim, num_vertices)).astype(np.float32)
