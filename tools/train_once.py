"""a few fused training steps (forward + backward) on the bench mesh: the short command the ncu captures of k_mlp_bwd / k_composite_bwd wrap"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tetra-nerf_b200")]
import torch
import bench
from tetranerf import cpp
from tetranerf.b200.render import FusedRenderer, RenderSettings

dev = torch.device("cuda:0")
syn = bench.synthetic()
V, C, field = bench.make_workload()
tr = cpp.TetrahedraTracer(dev)
dV, dC = torch.from_numpy(V).to(dev), torch.from_numpy(C).to(dev)
tr.load_tetrahedra(dV, dC)
fr = FusedRenderer(tr)
fr.set_field(torch.from_numpy(field).to(dev))
fr.set_weights(bench.mlp_params())
st = RenderSettings.tetra_nerf()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for i in range(steps):
    o, d = syn.camera_rays(R, seed=40 + i)
    jc, jf = torch.rand((R, 129), device=dev), torch.rand((R, 129), device=dev)
    out = fr.train_forward(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), st, jc, jf)
    g = (out["rgb"] - 0.5) / R
    gf, gp = fr.train_backward(g, None, len(V))
torch.cuda.synchronize()
print("train_once ok", float(gf.abs().max()))
