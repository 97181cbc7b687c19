"""API-op microbenchmark at BASELINE configs[1] sizes (4096 rays x 257 fine samples, 45k vertices, C=64):
interpolate_values forward and backward, backward with scalar atomics (reference formulation) vs the row-major
vector-reduction path.  Writes gpurun_out/r1_ops_bench.json."""
import json, os, sys
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, R_ + "/tetra-nerf_b200"]
import numpy as np, torch
from tetranerf.utils.extension import tetranerf_cpp_extension as ext

dev = torch.device("cuda:0")
V, Cdim, N = 45000, 64, 4096 * 257
rng = np.random.default_rng(0)
vi = torch.from_numpy(rng.integers(0, V, (N, 4)).astype(np.int32)).to(dev)
# samples along a ray stay in one tetrahedron for a few samples: repeat vertex sets in short runs (as the real workload)
run = torch.from_numpy(np.repeat(rng.integers(0, V, (N // 2 + 1, 4)), 2, axis=0)[:N].astype(np.int32)).to(dev)
w = torch.rand((N, 3), device=dev) / 4
field = torch.randn((Cdim, V), device=dev)
gin = torch.randn((N, Cdim), device=dev)
peak = json.load(open(R_ + "/MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists(R_ + "/MEASURED_PEAKS.json") else 6650.0


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))


res = {"N": N, "V": V, "C": Cdim, "hbm_peak_gbs": peak}
for name, idx in (("random vertices", vi), ("runs of 2 samples per tetrahedron", run)):
    r = {}
    r["forward_ms"] = timeit(lambda: ext.interpolate_values(idx, w, field))
    gf = torch.empty((Cdim, V), device=dev); scratch = torch.empty((V, Cdim), device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    def bwd(sc):
        ext._check(ext._lib.tn_interpolate_values_backward(0, 4, N, Cdim, V, idx.data_ptr(), w.data_ptr(), gin.data_ptr(), gf.data_ptr(), sc, s))
    r["backward_scalar_atomics_ms"] = timeit(lambda: bwd(None))
    a = gf.clone()
    r["backward_vector_reductions_ms"] = timeit(lambda: bwd(scratch.data_ptr()))
    r["max_abs_diff_between_paths"] = float((gf - a).abs().max().item())
    fwd_bytes = N * (16 + 12 + 4 * Cdim) + 4 * Cdim * V   # ids + weights + output + the table once
    bwd_bytes = N * (16 + 12 + 4 * Cdim) + 4 * Cdim * V
    r["forward_GBs_algorithmic"] = fwd_bytes / r["forward_ms"] / 1e6
    r["backward_GBs_algorithmic"] = bwd_bytes / r["backward_vector_reductions_ms"] / 1e6
    r["backward_frac_of_hbm_peak"] = r["backward_GBs_algorithmic"] / peak
    r["forward_frac_of_hbm_peak"] = r["forward_GBs_algorithmic"] / peak
    res[name] = {k: round(v, 4) for k, v in r.items()}
    print(name, res[name], flush=True)
# ---- find_visited_cells at configs[1] sizes: 4096 rays x 257 samples against a real trace of the 302k-tetrahedra mesh ----
import bench
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn
Vm, Cm, _ = bench.make_workload()
tr = cpp.TetrahedraTracer(dev); tr.load_tetrahedra(torch.from_numpy(Vm).to(dev), torch.from_numpy(Cm).to(dev))
o, d = syn.camera_rays(4096, seed=3); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
out = tr.trace_rays(o, d, 512)
hd = out["hit_distances"]; nv = out["num_visited_cells"].long().clamp_min(1)
near = hd[:, 0, 0:1]; far = torch.gather(hd[..., 1], 1, (nv - 1)[:, None])
dist = near + (far - near) * (torch.arange(257, device=dev)[None] + 0.5) / 257
fv = lambda: tr.find_visited_cells(out["num_visited_cells"], out["visited_cells"], out["barycentric_coordinates"], out["hit_distances"], out["vertex_indices"], dist)
res["find_visited_cells_ms (4096 x 257)"] = round(timeit(fv), 4)
res["trace_rays_dense_ms (4096, M=512)"] = round(timeit(lambda: tr.trace_rays(o, d, 512)), 4)
print({k: v for k, v in res.items() if "ms (" in k}, flush=True)
os.makedirs(R_ + "/gpurun_out", exist_ok=True)
json.dump(res, open(R_ + "/gpurun_out/r1_ops_bench.json", "w"), indent=1)
