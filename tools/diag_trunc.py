"""diagnostic: 1M-tet mesh, M=256 truncation, per-implementation mismatch report vs the oracle"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tetra-nerf_b200"), str(ROOT / "tests")]
import numpy as np, torch
from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn
import test_gpu_largemesh as T

V, C = syn.delaunay_mesh(150_000, seed=0)
om, tr = orc.OracleMesh(V, C), T._tracer(V, C)
o1, d1 = syn.camera_rays(384, seed=21); o2, d2 = syn.sphere_rays(384, seed=22); o3, d3 = T.diagonal_rays(256, seed=23)
o, d = np.concatenate([o1, o2, o3]), np.concatenate([d1, d2, d3])
M = 256
ref = om.trace_rays(o, d, M)
full = om.trace_rays(o, d, 1024)
for impl in T.IMPLS:
    g = T._gpu(tr, impl, o, d, M)
    stats = tr.trace_stats()
    bad = np.nonzero(g["num_visited_cells"] != ref["num_visited_cells"])[0]
    badc = np.nonzero((g["visited_cells"] != ref["visited_cells"]).any(1))[0]
    print(impl, "stats", stats, "rays with different num:", len(bad), " different cells:", len(badc))
    for r in list(badc[:6]):
        j = int(np.nonzero(g["visited_cells"][r] != ref["visited_cells"][r])[0][0])
        print(f"  ray {r}: num gpu {g['num_visited_cells'][r]} ref {ref['num_visited_cells'][r]} full(M=1024) {full['num_visited_cells'][r]} first diff at {j}:",
              g["visited_cells"][r, max(0, j - 2): j + 3], ref["visited_cells"][r, max(0, j - 2): j + 3],
              " t gpu", g["hit_distances"][r, max(0, j - 1): j + 2].ravel(), " t ref", ref["hit_distances"][r, max(0, j - 1): j + 2].ravel())
