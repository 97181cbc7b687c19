"""Second, independent oracle -- TEST INFRASTRUCTURE ONLY.

Per-tetrahedron plane clipping in float64: never looks at faces, face ids, sorting or pairing, so it
catches restatement bugs in tetra_oracle.cpp's gather/sort/post-process chain
(src/optix/optix_trace_rays.cu:78-331).  For a generic ray the reference emits exactly the tetrahedra
whose entry AND exit face are hit with t > 0 and whose crossing is >= 1e-6 long, ordered by entry t.
"""
import numpy as np


def tet_intervals(vertices, cells, o, d):
    """-> (tet ids sorted by t_in, t_in, t_out) for one ray, float64."""
    v = vertices.astype(np.float64)[cells.astype(np.int64)]  # [T,4,3]
    o = o.astype(np.float64)
    d = d.astype(np.float64)
    T = len(cells)
    t_in = np.full(T, -np.inf)
    t_out = np.full(T, np.inf)
    ok = np.ones(T, bool)
    for k in range(4):
        a, b, c = v[:, (k + 1) % 4], v[:, (k + 2) % 4], v[:, (k + 3) % 4]
        n = np.cross(b - a, c - a)
        flip = np.einsum("ij,ij->i", n, v[:, k] - a) > 0  # make n point away from the opposite vertex
        n[flip] *= -1
        s = np.einsum("ij,ij->i", n, o[None] - a)
        r = n @ d
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = -s / r
        ent = r < 0
        ext = r > 0
        par = r == 0
        t_in = np.where(ent, np.maximum(t_in, tt), t_in)
        t_out = np.where(ext, np.minimum(t_out, tt), t_out)
        ok &= ~(par & (s > 0))
    hit = ok & (t_in < t_out) & (t_in > 0)
    ids = np.nonzero(hit)[0]
    order = np.argsort(t_in[ids], kind="stable")
    ids = ids[order]
    return ids, t_in[ids], t_out[ids]
