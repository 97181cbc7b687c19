"""BASELINE.json configs[4]: synthetic ~1M-tetrahedra mesh, ray-batch sweep 1k..256k, traversal-only achieved GB/s vs the
HBM roofline.  Algorithmic bytes: B = 28 R + 52 sum(K) + 12 V + 16 T + 20 F.  Writes profiles/r1_traversal_sweep.json."""
import json, os, sys, time
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, R_ + "/tetra-nerf_b200"]
import numpy as np, torch
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn

dev = torch.device("cuda:0")
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000
t0 = time.time(); V, C = syn.delaunay_mesh(npts, seed=0); print("delaunay", C.shape, round(time.time() - t0, 1), "s", flush=True)
tr = cpp.TetrahedraTracer(dev)
dV, dC = torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev)
t0 = time.time(); tr.load_tetrahedra(dV, dC); torch.cuda.synchronize(); t_load = time.time() - t0
F = tr.num_faces()
peak = json.load(open(R_ + "/MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists(R_ + "/MEASURED_PEAKS.json") else 6650.0
M = 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"mesh": {"points": npts, "tetrahedra": int(len(C)), "faces": int(F), "load_tetrahedra_s": round(t_load, 3)}, "M": M, "hbm_peak_gbs": peak, "rows": []}
for gen_name, gen in (("camera", syn.camera_rays), ("sphere", syn.sphere_rays)):
    for R in (1024, 4096, 16384, 65536, 262144):
        o, d = gen(R, seed=9); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
        out = {"num_visited_cells": torch.empty((R,), dtype=torch.int32, device=dev), "visited_cells": torch.empty((R, M), dtype=torch.int32, device=dev),
               "barycentric_coordinates": torch.empty((R, M, 2, 3), device=dev), "hit_distances": torch.empty((R, M, 2), device=dev),
               "vertex_indices": torch.empty((R, M, 4), dtype=torch.int32, device=dev)}
        for _ in range(3): tr.trace_rays_into(o, d, M, out)
        torch.cuda.synchronize()
        ts = []
        for k in range(5):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.trace_rays_into(o, d, M, out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        sumK = int(out["num_visited_cells"].sum().item())
        B = 28 * R + 52 * sumK + 12 * len(V) + 16 * len(C) + 20 * F
        row = {"rays": gen_name, "R": R, "ms": round(ms, 4), "mean_K": round(sumK / R, 1), "max_K": int(out["num_visited_cells"].max().item()),
               "rays_per_s": round(R / ms * 1e3), "algorithmic_MB": round(B / 1e6, 1), "achieved_GBs": round(B / ms / 1e6, 1), "frac_of_hbm_peak": round(B / ms / 1e6 / peak, 4),
               "impl": "walk" if R >= 10240 else "bvh"}
        print(row, flush=True)
        res["rows"].append(row)
        del out
os.makedirs(R_ + "/gpurun_out", exist_ok=True)
json.dump(res, open(R_ + "/gpurun_out/r1_traversal_sweep.json", "w"), indent=1)
