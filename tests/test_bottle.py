"""The reference's only real geometry (tests/assets/bottle.ply -> tests/golden/bottle_mesh.npz, see make_bottle.py) with the
camera of its own test (tests/test_tetrahedra_tracer.py:24-59, generate_rays(64, 64)): surface-sampled points give slivers and
near-coplanar faces that uniform clouds never do.  CPU: the oracle against the float64 plane-clipping oracle and the reference's
geometric invariant (:204-207).  GPU: every implementation of trace_rays / find_visited_cells bit-equal to the oracle."""
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import intervals
from oracle import oracle as orc

GOLD = Path(__file__).resolve().parent / "golden" / "bottle_mesh.npz"


@pytest.fixture(scope="module")
def bottle():
    z = np.load(GOLD)
    return np.ascontiguousarray(z["vertices"], dtype=np.float32), np.ascontiguousarray(z["cells"], dtype=np.int32)


def reference_camera_rays(width=64, height=64):
    """restates generate_rays of the reference test: eye (0,1,0) looking at the origin, up z, fovY 45 degrees"""
    eye, lookat, up = np.array([0.0, 1.0, 0.0]), np.zeros(3), np.array([0.0, 0.0, 1.0])
    W = lookat - eye
    U = np.cross(W, up); U /= np.linalg.norm(U)
    Vv = np.cross(U, W); Vv /= np.linalg.norm(Vv)
    vlen = np.linalg.norm(W) * math.tan(0.5 * 45.0 * math.pi / 180.0)
    Vv *= vlen
    U *= vlen * (width / height)
    gx, gy = np.meshgrid(np.linspace(0, 1, width), np.linspace(0, 1, height), indexing="ij")
    dxy = 2.0 * np.stack([gy, gx], -1).reshape(-1, 2) - 1.0
    dirs = dxy[:, :1] * U[None] + dxy[:, 1:] * Vv[None] + W[None]
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    return np.repeat(eye[None], len(dirs), 0).astype(np.float32), dirs.astype(np.float32)


def test_bottle_oracle_against_plane_clipping(bottle):
    V, C = bottle
    assert len(V) > 2000 and len(C) > 10000
    m = orc.OracleMesh(V, C)
    o, d = reference_camera_rays()
    a = m.trace_rays(o, d, 256)
    hit_rays = np.nonzero(a["num_visited_cells"] > 0)[0]
    assert len(hit_rays) > 100
    total = missing = 0
    for i in hit_rays:
        ids, ti, to = intervals.tet_intervals(V, C, o[i], d[i])
        fin = np.isfinite(ti) & np.isfinite(to)  # zero-volume tetrahedra (coplanar surface samples) have no plane-clipping interval
        ids, ti, to = ids[fin], ti[fin], to[fin]
        n = a["num_visited_cells"][i]
        got = a["visited_cells"][i, :n].tolist()
        total += len(ids)
        it = iter(range(len(ids)))
        pos = []
        for g in got:
            for k in it:
                if ids[k] == g:
                    pos.append(k)
                    break
            else:
                pytest.fail(f"ray {i}: tetrahedron {g} not in order in the interval list")
        sliver = (to - ti) < 4e-6
        miss = sorted(set(range(len(ids))) - set(pos))
        missing += len(miss)
        for k in miss:
            assert sliver[max(0, k - 1): k + 2].any(), f"ray {i}: non-sliver tetrahedron {ids[k]} (len {to[k] - ti[k]:.3g}) missing"
        np.testing.assert_allclose(a["hit_distances"][i, :n, 0], ti[pos], atol=5e-5)
        np.testing.assert_allclose(a["hit_distances"][i, :n, 1], to[pos], atol=5e-5)
    print(f"bottle: {len(hit_rays)} rays hit, {total} intervals checked, {missing} sliver-adjacent records missing")
    # on surface-sampled geometry ~6 % of the plane-clipping intervals are slivers (or touch one) that the reference's eps = 1e-6
    # pairing drops (optix_trace_rays.cu:188-257); on uniform clouds it is 0.2 % (test_oracle.py).  Every missing one was checked above.
    assert missing <= 0.10 * total


def test_bottle_reference_invariant(bottle):
    """tests/test_tetrahedra_tracer.py:95-207: samples linspace(0.9, 1.1, 300); the matched points, rebuilt from the vertices and
    the interpolated barycentrics, lie on the ray (the reference allows |cos - 1| <= 0.05; here the point itself to 1e-5)"""
    V, C = bottle
    m = orc.OracleMesh(V, C)
    o, d = reference_camera_rays()
    tr = m.trace_rays(o, d, 256)
    dist = np.ascontiguousarray(np.broadcast_to(np.linspace(0.90, 1.1, 300, dtype=np.float32), (len(o), 300)))
    mt = orc.find_visited_cells(tr["num_visited_cells"], tr["visited_cells"], tr["barycentric_coordinates"], tr["hit_distances"],
                                tr["vertex_indices"], dist)
    mask = mt["mask"]
    assert mask.sum() > 1000
    b = mt["barycentric_coordinates"][mask].astype(np.float64)
    w = np.concatenate([1 - b.sum(-1, keepdims=True), b], -1)
    pts = (V.astype(np.float64)[mt["vertex_indices"][mask]] * w[..., None]).sum(-2)
    rr, ss = np.nonzero(mask)
    want = o[rr].astype(np.float64) + dist[rr, ss][:, None].astype(np.float64) * d[rr].astype(np.float64)
    err = np.abs(pts - want).max()
    print("bottle: matched samples", int(mask.sum()), "max |point - (o + t d)| =", err)
    assert err < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("impl", ["walk", "walk_solo", "walk_quad", "walk_quad_pf", "bvh"])
def test_bottle_gpu_equals_oracle(bottle, impl):
    from tetranerf import cpp

    dev = torch.device("cuda:0")
    V, C = bottle
    tr = cpp.TetrahedraTracer(dev)
    tr.load_tetrahedra(torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev))
    from conftest import force_trace_impl

    force_trace_impl(tr, impl)
    o, d = reference_camera_rays()
    o2 = (o + np.random.default_rng(0).normal(0, 0.3, o.shape)).astype(np.float32)  # a second, incoherent bundle aimed at the bottle
    d2 = -o2 + np.random.default_rng(1).normal(0, 0.03, o.shape); d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    oo, dd = np.concatenate([o, o2]), np.concatenate([d, d2])
    out = tr.trace_rays(torch.from_numpy(oo).to(dev), torch.from_numpy(dd).to(dev), 256)
    tr.synchronize()
    walkable, n_exact = tr.trace_stats()
    ref = orc.OracleMesh(V, C).trace_rays(oo, dd, 256)
    hit = int((ref["num_visited_cells"] > 0).sum())
    print(f"bottle/{impl}: walkable={walkable}, {hit} of {len(oo)} rays hit, {n_exact} rays took the exact stage ({100.0 * n_exact / max(hit, 1):.2f} % of hits)")
    for k, v in ref.items():
        assert np.array_equal(out[k].cpu().numpy().view(np.uint32), v.view(np.uint32)), f"bottle/{impl}: {k} differs from the oracle"
    dist = torch.linspace(0.90, 1.1, 300, device=dev).expand(len(oo), 300).contiguous()
    mt = tr.find_visited_cells(out["num_visited_cells"], out["visited_cells"], out["barycentric_coordinates"], out["hit_distances"],
                               out["vertex_indices"], dist)
    mr = orc.find_visited_cells(ref["num_visited_cells"], ref["visited_cells"], ref["barycentric_coordinates"], ref["hit_distances"],
                                ref["vertex_indices"], dist.cpu().numpy())
    assert np.array_equal(mt["mask"].cpu().numpy(), mr["mask"]) and np.array_equal(mt["cell_indices"].cpu().numpy(), mr["cell_indices"])
    assert np.array_equal(mt["barycentric_coordinates"].cpu().numpy().view(np.uint32), mr["barycentric_coordinates"].view(np.uint32))
