"""How many rays of the bench workload the quad walk hands to the exact stage (and how many of those need the all-hits gather
rather than only the sort/pairing of the walk's own keys).  Run under `ncu --metrics gpu__time_duration.sum` to split the trace
time into k_walk_quad and the exact stage (k_trace<0>)."""
import ctypes, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tetra-nerf_b200"]
import torch, bench
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn
from tetranerf.utils.extension import tetranerf_cpp_extension as ext

dev = torch.device("cuda:0")
V, C, field = bench.make_workload()
tr = cpp.TetrahedraTracer(dev); tr.load_tetrahedra(torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev))
tr.set_walk_quad_range(0, 2**32 - 1)
res = {}
for n in (4096, 8192):
    o, d = syn.camera_rays(n, seed=3); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    for _ in range(3): out = tr.trace_rays(o, d, 512)
    walkable, listed = tr.trace_stats()
    gather = int(ext._lib.tn_debug_last_exact_count())
    nv = out["num_visited_cells"]
    res[n] = {"listed": listed, "all_hits_gather": gather, "pairing_only": listed - gather,
              "records_mean": float(nv.float().mean()), "records_max": int(nv.max())}
# in-pipeline timing (warm L2, as inside tn_render): CUDA events around trace_rays
for n in (4096, 8192):
    o, d = syn.camera_rays(n, seed=3); o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    ts = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr.trace_rays(o, d, 512); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[n]["trace_rays_ms_median"] = round(sorted(ts)[len(ts) // 2], 4)
print(json.dumps(res))
