"""CPU tests of the oracle itself: pinned against the float64 known-answer crossing list of SURVEY.md §8c,
against an independent plane-clipping oracle, against brute force, and against the invariants the
reference's own tests assert (tests/test_tetrahedra_tracer.py:204-207, :410-416, :442-453)."""
import numpy as np
import pytest
import torch

from oracle import intervals
from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn


def test_unique_faces_cube(cube_mesh):
    V, C = cube_mesh
    m = orc.OracleMesh(V, C)
    tri, tt = m.faces()
    assert m.num_faces == 30  # 12 hull faces... (F=30)
    # first tetra (0,1,2,8): faces j=0..3 in the rotation of tetrahedra_tracer.cpp:54-57
    assert tri[:4].tolist() == [[1, 2, 8], [2, 8, 0], [8, 0, 1], [0, 1, 2]]
    assert (tt[:, 0] != 0xFFFFFFFF).all()
    hull = (tt[:, 1] == 0xFFFFFFFF).sum()
    assert hull == 12
    # every face's first owner is the lower-indexed tetrahedron
    inner = tt[:, 1] != 0xFFFFFFFF
    assert (tt[inner, 0] < tt[inner, 1]).all()


def test_face_shared_by_three_raises():
    V = np.random.default_rng(0).random((6, 3)).astype(np.float32)
    C = np.array([[0, 1, 2, 3], [0, 1, 2, 4], [0, 1, 2, 5]], np.int32)
    with pytest.raises(RuntimeError, match="more than two"):
        orc.OracleMesh(V, C)


def test_known_answer_cube_ray(cube_mesh):
    """float64 geometry with margins >= 0.07 from every edge."""
    V, C = cube_mesh
    m = orc.OracleMesh(V, C)
    d = np.array([1, 0.13, 0.29])
    d = (d / np.linalg.norm(d)).astype(np.float32)
    for accel in (True, False):
        out = m.trace_rays(np.array([[-0.05, 0.07, 0.21]], np.float32), d[None], 16, accel=accel)
        assert out["num_visited_cells"][0] == 4
        assert out["visited_cells"][0, :4].tolist() == [4, 2, 3, 10]
        assert (out["visited_cells"][0, 4:] == -1).all() and (out["vertex_indices"][0, 4:] == -1).all()
        ts = [0.052464, 0.144729, 0.683256, 0.910000, 1.101750]
        np.testing.assert_allclose(out["hit_distances"][0, :4, 0], ts[:4], atol=2e-6)
        np.testing.assert_allclose(out["hit_distances"][0, :4, 1], ts[1:], atol=2e-6)
        assert out["vertex_indices"][0, 0].tolist() == [8, 0, 2, 4]
        np.testing.assert_allclose(out["barycentric_coordinates"][0, 0, 0], [0.6990, 0.0765, 0.2245], atol=1e-4)
        np.testing.assert_allclose(out["barycentric_coordinates"][0, 0, 1], [0.6621, 0.0, 0.1621], atol=1e-4)


def test_reference_degenerate_ray_is_stable(cube_mesh):
    """tests/test_tetrahedra_tracer.py:255-256: passes through two mesh edges and inside a face; the reference
    asserts nothing.  Robustness only: no crash, sane bounds, deterministic."""
    V, C = cube_mesh
    m = orc.OracleMesh(V, C)
    a = m.trace_rays(np.array([[-0.05, 0.05, 0.05]], np.float32), np.array([[1, 0, 0]], np.float32), 16)
    b = m.trace_rays(np.array([[-0.05, 0.05, 0.05]], np.float32), np.array([[1, 0, 0]], np.float32), 16, accel=False)
    for k in a:
        assert np.array_equal(a[k], b[k])
    n = a["num_visited_cells"][0]
    assert 0 <= n <= 15 and (a["visited_cells"][0, :n] < 12).all()


def test_power_of_two_check(cube_mesh):
    m = orc.OracleMesh(*cube_mesh)
    with pytest.raises(RuntimeError, match="power of 2"):
        m.trace_rays(np.zeros((1, 3), np.float32), np.ones((1, 3), np.float32), 12)


@pytest.mark.parametrize("gen", [syn.camera_rays, syn.sphere_rays])
def test_bvh_equals_brute_force(small_mesh, gen):
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = gen(200)
    a = m.trace_rays(o, d, 256)
    b = m.trace_rays(o, d, 256, accel=False)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["num_visited_cells"].max() > 40


@pytest.mark.parametrize("gen", [syn.camera_rays, syn.sphere_rays])
def test_against_plane_clipping_oracle(small_mesh, gen):
    """Independent float64 per-tetrahedron interval test.  The face-pairing result must be a subsequence of
    the interval list; anything missing is a sliver (< 4e-6 long, the reference's eps=1e-6 is ~4 ulp of t)
    or the neighbour of one (tie order at equal fp32 t can break a pairing)."""
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = gen(150)
    a = m.trace_rays(o, d, 512)
    total = missing = 0
    for i in range(len(o)):
        ids, ti, to = intervals.tet_intervals(V, C, o[i], d[i])
        n = a["num_visited_cells"][i]
        got = a["visited_cells"][i, :n].tolist()
        total += len(ids)
        it = iter(range(len(ids)))
        pos = []
        for g in got:  # subsequence check
            for k in it:
                if ids[k] == g:
                    pos.append(k)
                    break
            else:
                pytest.fail(f"ray {i}: tetra {g} not in order in the interval list")
        sliver = (to - ti) < 4e-6
        miss = sorted(set(range(len(ids))) - set(pos))
        missing += len(miss)
        for k in miss:
            near = sliver[max(0, k - 1) : k + 2].any()
            assert near, f"ray {i}: non-sliver tetra {ids[k]} (len {to[k]-ti[k]:.3g}) missing"
        np.testing.assert_allclose(a["hit_distances"][i, :n, 0], ti[pos], atol=5e-5)  # grazing faces: t is ill-conditioned in fp32
        np.testing.assert_allclose(a["hit_distances"][i, :n, 1], to[pos], atol=5e-5)
    assert missing <= 0.002 * total + 2


def test_barycentrics_reconstruct_points(small_mesh):
    """Entry/exit barycentrics + vertex_indices reproduce o + t d (the geometric invariant of
    tests/test_tetrahedra_tracer.py:204-207, here to 1e-5 absolute)."""
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = syn.camera_rays(64)
    a = m.trace_rays(o, d, 256)
    for i in range(len(o)):
        n = a["num_visited_cells"][i]
        vi = a["vertex_indices"][i, :n]
        for side in (0, 1):
            b = a["barycentric_coordinates"][i, :n, side]
            w = np.concatenate([1 - b.sum(-1, keepdims=True), b], -1)
            p = (V[vi] * w[..., None]).sum(1)
            t = a["hit_distances"][i, :n, side]
            np.testing.assert_allclose(p, o[i] + t[:, None] * d[i], atol=2e-5)
        # the 4 vertices are exactly the visited cell's vertices
        assert (np.sort(vi, 1) == np.sort(C[a["visited_cells"][i, :n]], 1)).all()


def test_hit_cap_keeps_nearest(small_mesh):
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = syn.camera_rays(32)
    full = m.trace_rays(o, d, 512)
    cap = m.trace_rays(o, d, 16)
    for i in range(len(o)):
        n = cap["num_visited_cells"][i]
        assert n <= 14
        assert np.array_equal(cap["visited_cells"][i, :n], full["visited_cells"][i, :n])


def test_matcher_and_interpolation_identities(small_mesh):
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = syn.camera_rays(48)
    tr = m.trace_rays(o, d, 256)
    S = 100
    dist = np.linspace(1.0, 3.0, S, dtype=np.float32)[None].repeat(len(o), 0)
    mc = orc.find_visited_cells(tr["num_visited_cells"], tr["visited_cells"], tr["barycentric_coordinates"], tr["hit_distances"],
                                tr["vertex_indices"], dist)
    assert mc["mask"].any() and not mc["mask"].all()
    assert (mc["cell_indices"][~mc["mask"]] == -1).all() and (mc["vertex_indices"][~mc["mask"]] == -1).all()
    # matched samples: interpolated position lies on the ray (reference invariant :204-207)
    b = mc["barycentric_coordinates"][mc["mask"]]
    vi = mc["vertex_indices"][mc["mask"]]
    w = np.concatenate([1 - b.sum(-1, keepdims=True), b], -1)
    p = (V[vi] * w[..., None]).sum(1)
    ii, jj = np.nonzero(mc["mask"])
    np.testing.assert_allclose(p, o[ii] + dist[ii, jj][:, None] * d[ii], atol=3e-5)
    # interpolation == einsum over safe-gathered field (reference test :410-416) and its gradient (:442-453)
    field = torch.from_numpy(syn.random_field(len(V), 8)).requires_grad_(True)
    out = orc.interpolate_values(mc["vertex_indices"], mc["barycentric_coordinates"], field.detach().numpy())
    vi_t = torch.from_numpy(mc["vertex_indices"]).long()
    safe = vi_t.clamp_min(0)
    g = torch.where((vi_t >= 0)[None], field[:, safe], torch.zeros(()))  # [C,R,S,4]
    wt = torch.from_numpy(np.concatenate([1 - mc["barycentric_coordinates"].sum(-1, keepdims=True), mc["barycentric_coordinates"]], -1))
    gt = torch.einsum("jrbi,rbi->rbj", g, wt)
    np.testing.assert_allclose(out, gt.detach().numpy(), rtol=1.3e-6, atol=1e-5)
    gt.sum().backward()
    gb = orc.interpolate_values_backward(mc["vertex_indices"], mc["barycentric_coordinates"], tuple(field.shape), np.ones_like(out))
    np.testing.assert_allclose(gb, field.grad.numpy(), rtol=1e-5, atol=1e-4)


def test_find_tetrahedra_reference_golden(cube_mesh):
    """The only value-pinning tracer test of the reference: tests/test_tetrahedra_tracer.py:270-344."""
    V, C = cube_mesh
    m = orc.OracleMesh(V, C)

    def mix(*a):
        return sum(V[i] * w for i, w in zip(a[::2], a[1::2]))

    pts = np.stack([mix(0, 0.23, 1, 0.27, 2, 0.21, 8, 0.29), mix(2, 0.23, 4, 0.24, 6, 0.26, 8, 0.27), mix(3, 0.39, 5, 0.41, 7, 0.09, 8, 0.11)]).astype(np.float32)
    out = m.find_tetrahedra(pts)
    assert out["tetrahedra"].tolist() == [0, 5, 11]
    gt_coords = np.array([[0.23, 0.27, 0.21, 0.29], [0.23, 0.24, 0.26, 0.27], [0.39, 0.41, 0.09, 0.11]], np.float32)
    gt_idx = np.array([[0, 1, 2, 8], [2, 4, 6, 8], [3, 5, 7, 8]])
    bc = out["barycentric_coordinates"]
    bc = np.concatenate([1 - bc.sum(-1, keepdims=True), bc], -1)
    for i in range(3):
        order = np.argsort(out["vertex_indices"][i])
        assert (out["vertex_indices"][i][order] == gt_idx[i]).all()
        np.testing.assert_allclose(bc[i][order], gt_coords[i], rtol=1.3e-6, atol=1e-5)


def test_oracle_render_sanity(small_mesh):
    V, C = small_mesh
    m = orc.OracleMesh(V, C)
    o, d = syn.camera_rays(24)
    o[0] = [5, 5, 5]; d[0] = [1, 0, 0]  # misses the mesh
    field = torch.from_numpy(syn.random_field(len(V)))
    params = orc.init_mlp_params(0)
    for cfg in (orc.RenderConfig.tetra_nerf(), orc.RenderConfig(num_samples=32, num_fine_samples=32)):
        out = orc.render(m, field, params, o, d, cfg, return_aux=True)
        assert out["rgb"].shape == (24, 3) and not bool(out["ray_mask"][0])
        assert torch.all(out["rgb"][0] == 1) and float(out["depth"][0]) == cfg.far_plane and float(out["accumulation"][0]) == 0
        assert torch.isfinite(out["rgb"]).all() and (out["rgb"] >= 0).all() and (out["rgb"] <= 1).all()
        assert (out["accumulation"] <= 1 + 1e-5).all()
        e = out["aux"]["fine_euclid"]
        assert e.shape[1] == cfg.num_samples + cfg.num_fine_samples + 2 and (e[:, 1:] >= e[:, :-1]).all()
