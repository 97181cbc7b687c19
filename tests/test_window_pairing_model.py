"""CPU: the windowing argument of tn_trace.cu:post_process_windows, checked on a model.

The reference's post_process_tetrahedra (optix_trace_rays.cu:110-266) is two serial loops over the sorted face hits of a ray.  The
CUDA exact stage runs their bodies only inside windows around eps-ties (and wherever the swaps of the pairing loop reach) and treats
every other position as "pairs with its successor".  This file restates both forms in Python -- `literal` = the two loops over the
whole array, `windowed` = the control flow of post_process_windows with the same two bodies -- and asserts that they produce the same
emitted pairs and the same final array on (a) the sorted raw hits of real rays through a sliver mesh (oracle trace_rays_triangles +
the oracle's face table) and (b) random adversarial sequences (clusters of ties, shared / unshared tetrahedra, hull faces).  The GPU
tests (test_gpu_slivers.py, test_gpu_trace.py) then hold the CUDA code against the oracle bit for bit."""
import copy

import numpy as np

EMPTY = 0xFFFFFFFF
EPS = np.float32(1e-6)
f32 = np.float32


def common(a, b):  # optix_trace_rays.cu:22-37
    return a[0] == b[0] or a[0] == b[1] or a[1] == b[0] or a[1] == b[1]


class Hits:
    def __init__(self, t, face, tts):
        self.t = [f32(x) for x in t]
        self.face = [int(x) for x in face]
        self.mark = [False] * len(self.t)
        self.tts = [tuple(int(v) for v in x) for x in tts]
        self.n = len(self.t)

    def state(self):
        return (self.t, self.face, self.mark, self.tts)


def dedupe_one(A, j):  # body of :124-159
    dn, clear_self, off, n = A.t[j], False, 1, A.n
    while j + off < n and (A.face[j + off] == EMPTY or abs(f32(A.t[j + off] - dn)) < EPS):
        if A.face[j + off] != EMPTY and common(A.tts[j], A.tts[j + off]):
            if A.face[j] != A.face[j + off]:
                clear_self = True
            if A.mark[j + off]:
                A.face[j + off] = EMPTY
            else:
                A.mark[j + off] = True
        off += 1
    if clear_self and A.mark[j]:
        A.face[j] = EMPTY
    A.mark[j] = False


def pair_one(A, j):  # body of :161-258; returns (emitted, largest index written)
    orig, dn, ro, off, n, wr = A.tts[j], A.t[j], 1, 1, A.n, 0
    while j + off < n and (ro < 3 or A.face[j + off] == EMPTY or abs(f32(A.t[j + off] - dn)) < EPS):
        if A.face[j + off] == EMPTY:
            off += 1
            continue
        if common(orig, A.tts[j + off]):
            out = abs(f32(A.t[j] - A.t[j + off])) >= EPS
            if off > 1:
                for L in A.state():
                    L[j + off], L[j + 1] = L[j + 1], L[j + off]
                wr = j + off
            return bool(out), wr
        dn = A.t[j + off]
        ro += 1
        off += 1
    return False, wr


def literal(A):
    for j in range(A.n - 1):
        if A.face[j] != EMPTY:
            dedupe_one(A, j)
    return [j for j in range(A.n) if A.face[j] != EMPTY and pair_one(A, j)[0]]


def windowed(A):
    n = A.n
    link = [j + 1 < n and (bool(abs(f32(A.t[j + 1] - A.t[j])) < EPS) or not common(A.tts[j], A.tts[j + 1])) for j in range(n)]
    W = [any(link[i] for i in range(max(j - 3, 0), min(j + 2, n))) for j in range(n)]
    E = [False] * n

    def next_set(frm):
        while frm < n and not W[frm]:
            frm += 1
        return frm

    def run_end(ws):
        while ws + 1 < n and W[ws + 1]:
            ws += 1
        return ws

    ws = next_set(0)
    while ws < n:
        we = run_end(ws)
        for j in range(ws, min(we, n - 2) + 1):
            if A.face[j] != EMPTY:
                dedupe_one(A, j)
        ws = next_set(we + 1)
    ws = next_set(0)
    work = 0
    while ws < n:
        we = run_end(ws)
        lim, j = we, ws
        while j <= lim and j < n:
            em, wr = (False, 0) if A.face[j] == EMPTY else pair_one(A, j)
            lim = max(lim, wr)
            W[j] = True
            E[j] = em
            j += 1
            work += 1
        ws = next_set(lim + 1)
    return [j for j in range(n) if (E[j] if W[j] else j + 1 < n)], work


def _check(A):
    B = copy.deepcopy(A)
    e1 = literal(A)
    e2, work = windowed(B)
    assert e1 == e2
    assert A.state() == B.state()
    return work


def test_windowed_equals_literal_on_sliver_mesh_rays():
    from oracle import oracle as orc
    from tetranerf.b200 import synthetic as syn
    from test_gpu_slivers import sliver_mesh

    total = work = 0
    for jitter, per, seed in ((2e-6, 3, 5), (5e-7, 4, 6)):
        V, C = sliver_mesh(sites=300, per=per, jitter=jitter, seed=seed)
        m = orc.OracleMesh(V, C)
        _, tt = m.faces()
        o, d = syn.camera_rays(300, seed=4)
        r = m.trace_rays_triangles(o, d, 512)
        for i in range(len(o)):
            k = int(r["num_visited_triangles"][i])
            if k >= 2:
                faces = r["visited_triangles"][i, :k]
                work += _check(Hits(r["hit_distances"][i, :k], faces, tt[faces]))
                total += k
    assert total > 10000
    print(f"pairing bodies executed by the windowed form: {work} of {total} positions")


def test_windowed_equals_literal_on_adversarial_sequences():
    rng = np.random.default_rng(0)
    for case in range(3000):
        n = int(rng.integers(2, 60))
        # increasing t with clusters of (near-)ties
        steps = np.where(rng.random(n) < 0.35, rng.choice([0.0, 2.4e-7, 4.8e-7, 9.5e-7, 1.2e-6], n), rng.uniform(1e-4, 1e-2, n))
        t = np.cumsum(steps).astype(np.float32) + np.float32(1.0)
        # a walk through tetrahedra 0,1,2,...: face k separates tet k-1 and tet k; perturbed so that neighbours sometimes share nothing,
        # sometimes share through the second slot, and hull faces (EMPTY on one side) appear anywhere
        tts = []
        for k in range(n):
            a, b = k, k + 1
            u = rng.random()
            if u < 0.10:
                a, b = int(rng.integers(0, n + 2)), int(rng.integers(0, n + 2))
            elif u < 0.18:
                b = EMPTY
            elif u < 0.30:
                a, b = b, a
            tts.append((a, b))
        faces = rng.permutation(1000)[:n]
        if rng.random() < 0.3:  # local reorderings, as a sort by (t, face) produces among exact ties
            j = int(rng.integers(0, n - 1))
            tts[j], tts[j + 1] = tts[j + 1], tts[j]
        _check(Hits(t, faces, tts))
