"""BASELINE.json configs[4]: synthetic ~1M-tetrahedra mesh, ray-batch sweep 1k..256k PER GPU, traversal-only achieved GB/s vs the HBM
roofline, at 1/2/4/8 GPUs (torchrun: every rank traces its own batch of the same size = weak scaling, time = max over ranks).
Algorithmic bytes: B = 28 R + 52 sum(K) + 12 V + 16 T + 20 F per GPU.  Writes gpurun_out/r2_traversal_sweep_n<N>.json.

    python tools/traversal_sweep.py [points]                                              # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/traversal_sweep.py       # N GPUs
"""
import json, os, sys, time
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, R_ + "/tetra-nerf_b200"]
import numpy as np, torch
import torch.distributed as dist
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 150_000
t0 = time.time(); V, C = syn.delaunay_mesh(npts, seed=0)
if rank == 0: print("delaunay", C.shape, round(time.time() - t0, 1), "s", flush=True)
tr = cpp.TetrahedraTracer(dev)
dV, dC = torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev)
t0 = time.time(); tr.load_tetrahedra(dV, dC); torch.cuda.synchronize(); t_load = time.time() - t0
F = tr.num_faces()
peak = json.load(open(R_ + "/MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists(R_ + "/MEASURED_PEAKS.json") else 6650.0
M = 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"n_gpus": world, "mesh": {"points": npts, "tetrahedra": int(len(C)), "faces": int(F), "load_tetrahedra_s": round(t_load, 3)}, "M": M, "hbm_peak_gbs": peak,
       "scaling": "weak: every GPU traces its own batch of R rays; time = max over ranks; rays_per_s and GB/s are whole-job", "rows": []}
for gen_name, gen in (("camera", syn.camera_rays), ("sphere", syn.sphere_rays)):
    for R in (1024, 4096, 16384, 65536, 262144):
        o, d = gen(R, seed=9 + 100 * rank); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
        out = {"num_visited_cells": torch.empty((R,), dtype=torch.int32, device=dev), "visited_cells": torch.empty((R, M), dtype=torch.int32, device=dev),
               "barycentric_coordinates": torch.empty((R, M, 2, 3), device=dev), "hit_distances": torch.empty((R, M, 2), device=dev),
               "vertex_indices": torch.empty((R, M, 4), dtype=torch.int32, device=dev)}
        for _ in range(3): tr.trace_rays_into(o, d, M, out)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        ts = []
        for k in range(5):
            flush.fill_(k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tr.trace_rays_into(o, d, M, out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts)), float(out["num_visited_cells"].sum().item())], dtype=torch.float64, device=dev)
        tmax, ksum = t[0:1].clone(), t[1:2].clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(ksum, op=dist.ReduceOp.SUM)
        ms, sumK = float(tmax.item()), int(ksum.item())
        B = 28 * R * world + 52 * sumK + world * (12 * len(V) + 16 * len(C) + 20 * F)
        impl = "bvh gather" if R < 3584 else ("walk, 8 rays/warp" if R < (1 << 20) else "walk, 32 rays/warp")
        row = {"rays": gen_name, "R_per_gpu": R, "ms": round(ms, 4), "mean_K": round(sumK / (R * world), 1), "rays_per_s": round(R * world / ms * 1e3),
               "algorithmic_MB": round(B / 1e6, 1), "achieved_GBs": round(B / ms / 1e6, 1), "frac_of_hbm_peak": round(B / ms / 1e6 / (peak * world), 4), "impl": impl}
        if rank == 0:
            print(row, flush=True)
            res["rows"].append(row)
        del out
if rank == 0:
    os.makedirs(R_ + "/gpurun_out", exist_ok=True)
    json.dump(res, open(R_ + f"/gpurun_out/r2_traversal_sweep_n{world}.json", "w"), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
