"""Generates tests/golden/bottle_mesh.npz from the reference's only real-geometry test asset -- run in the build container:

    python tests/golden/make_bottle.py            (reads /root/reference/tests/assets/bottle.ply; the GPU box has no /root/reference)

The reference's fixture (tests/test_tetrahedra_tracer.py:13-21) loads the PLY with trimesh and tetrahedralises its vertices with
CGAL (`cpp.triangulate`); neither is available here, so the binary PLY is parsed directly (2,549 surface vertices, x y z of 8
float32 properties) and tetrahedralised with scipy's Qhull Delaunay -- cell order/orientation differ from CGAL's, which is
irrelevant once the mesh is an input.  Surface-sampled points give slivers and near-coplanar faces that the uniform random
clouds of the other fixtures never produce.  The fixture stores vertices f32[V,3] (exact duplicates removed) and cells i32[T,4]."""
import sys
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/tests/assets/bottle.ply")
OUT = Path(__file__).resolve().parent / "bottle_mesh.npz"


def read_ply_vertices(path: Path) -> np.ndarray:
    raw = path.read_bytes()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    assert "format binary_little_endian 1.0" in header
    nv = int(next(h for h in header if h.startswith("element vertex")).split()[-1])
    props = []
    in_vertex = False
    for h in header:
        if h.startswith("element"):
            in_vertex = h.startswith("element vertex")
        elif in_vertex and h.startswith("property float"):
            props.append(h.split()[-1])
    assert props[:3] == ["x", "y", "z"]
    data = np.frombuffer(raw, dtype="<f4", count=nv * len(props), offset=end).reshape(nv, len(props))
    return np.ascontiguousarray(data[:, :3]).astype(np.float32)


def main():
    from scipy.spatial import Delaunay

    v = read_ply_vertices(SRC)
    v = np.unique(v, axis=0)  # exact duplicates (seams) would be left unreferenced by Qhull
    cells = Delaunay(v.astype(np.float64)).simplices.astype(np.int32)
    np.savez_compressed(OUT, vertices=v, cells=np.ascontiguousarray(cells))
    print(f"{OUT}: {len(v)} vertices, {len(cells)} tetrahedra, bbox {v.min(0)} .. {v.max(0)}")


if __name__ == "__main__":
    sys.exit(main())
