"""Generates tests/golden/*.npz -- run from the repo root:  python tests/golden/make_golden.py

The reference itself cannot run here (OptiX-only tracer, nerfstudio not installed), so the golden
vectors are (a) the float64 known-answer crossing list of the cube mesh derived in SURVEY.md §8c, stored in
cube_kat.json by hand, and (b) outputs of the pinned CPU oracle (oracle/) on a small mesh that is stored inside the
fixture (so it does not depend on the scipy/Qhull version): regression pins for every stage of the path."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tetra-nerf_b200")]
from oracle import oracle as orc  # noqa: E402
from tetranerf.b200 import synthetic as syn  # noqa: E402


def main():
    V, C = syn.delaunay_mesh(600, seed=42)
    mesh = orc.OracleMesh(V, C)
    tri, tt = mesh.faces()
    o1, d1 = syn.camera_rays(48, seed=7)
    o2, d2 = syn.sphere_rays(48, seed=8)
    o, d = np.concatenate([o1, o2]), np.concatenate([d1, d2])
    o[5] = [4, 4, 4]; d[5] = [1, 0, 0]  # a ray that misses
    tr = mesh.trace_rays(o, d, 128)
    tr16 = mesh.trace_rays(o, d, 16)    # hit cap active
    S = 40
    dist = np.sort(np.random.default_rng(1).uniform(0.9, 3.3, (len(o), S)).astype(np.float32), axis=1)
    mc = orc.find_visited_cells(tr["num_visited_cells"], tr["visited_cells"], tr["barycentric_coordinates"], tr["hit_distances"],
                                tr["vertex_indices"], dist)
    field = syn.random_field(len(V), 64, seed=3)
    feat = orc.interpolate_values(mc["vertex_indices"], mc["barycentric_coordinates"], field)
    params = orc.init_mlp_params(0)
    cfg = orc.RenderConfig(num_samples=32, num_fine_samples=32, use_biased_sampler=True, max_intersected_triangles=128)
    img = orc.render(mesh, torch.from_numpy(field), params, o, d, cfg)
    pts = (np.random.default_rng(2).random((40, 3)) * 1.2 - 0.1).astype(np.float32)
    ft = mesh.find_tetrahedra(pts)
    np.savez_compressed(
        Path(__file__).parent / "small_mesh_golden.npz",
        vertices=V, cells=C, tri=tri, tt=tt, origins=o, directions=d,
        **{f"trace_{k}": v for k, v in tr.items()}, **{f"trace16_{k}": v for k, v in tr16.items()},
        sample_distances=dist, **{f"match_{k}": v for k, v in mc.items()}, features_checksum=np.float64(feat.astype(np.float64).sum()),
        features_first=feat[0, :4], render_rgb=img["rgb"].numpy(), render_acc=img["accumulation"].numpy(), render_depth=img["depth"].numpy(),
        render_mask=img["ray_mask"].numpy(), find_points=pts, **{f"find_{k}": v for k, v in ft.items()},
    )
    print("wrote", Path(__file__).parent / "small_mesh_golden.npz")


if __name__ == "__main__":
    main()
