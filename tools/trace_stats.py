import os, sys, ctypes, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R, R + "/tetra-nerf_b200"]
import bench
from tetranerf import cpp
from tetranerf.b200 import synthetic as syn
dev = torch.device("cuda:0")
V, C, field = bench.make_workload()
tr = cpp.TetrahedraTracer(dev)
dV, dC = torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev)
tr.load_tetrahedra(dV, dC)
o, d = syn.camera_rays(4096, seed=5)
o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
for _ in range(3):
    out = tr.trace_rays(o, d, 512)
print("stats (walkable, listed):", tr.trace_stats())
from tetranerf.utils.extension import tetranerf_cpp_extension as ext
print("  of which need the all-hits gather:", ext._lib.tn_debug_last_exact_count())
# time pieces with events
import time
torch.cuda.synchronize()
for M in (512,):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = tr.trace_rays(o, d, M)
    e1.record(); torch.cuda.synchronize()
    print("trace_rays dense M=%d: %.3f ms/call" % (M, e0.elapsed_time(e1) / 10))
