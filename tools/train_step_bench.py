"""Training step through the drop-in TetrahedraNerf (BASELINE configs[2] shape, on the single-GPU bench mesh): 8192 rays,
`tetra-nerf` config, train mode, forward + backward + Adam step.  The training path is the reference's op sequence on this
repo's CUDA ops (trace_rays, find_visited_cells, interpolate_values fwd/bwd) with a torch fp32 MLP and autograd -- the fused
tcgen05 path is inference-only this round.  Writes gpurun_out/r1_train_step.json."""
import json, os, sys, time
R_ = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [R_, R_ + "/tetra-nerf_b200"]
import numpy as np, torch
import bench
from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn
from tetranerf.nerfstudio import model as M

dev = torch.device("cuda:0")
V, C, field = bench.make_workload()
cfg = M.TetrahedraNerfConfig(num_tetrahedra_vertices=len(V), num_tetrahedra_cells=len(C), num_samples=128, num_fine_samples=128, use_biased_sampler=True)
m = M.TetrahedraNerf(cfg)
sd = {"tetrahedra_vertices": torch.from_numpy(V), "tetrahedra_cells": torch.from_numpy(C), "tetrahedra_field": torch.from_numpy(field)}
sd.update(orc.init_mlp_params(0))
m.load_state_dict(sd, strict=False)
m = m.to(dev).train()
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batches = []
for i in range(8):
    o, d = syn.camera_rays(R, seed=50 + i)
    batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.rand((R, 3), device=dev)))


def step(i, phases=None):
    o, d, img = batches[i % len(batches)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    out = m(M.RayBundle(origins=o, directions=d))
    ev[1].record()
    loss = m.get_loss_dict(out, {"image": img})["rgb_loss"]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    if phases is not None:
        torch.cuda.synchronize()
        phases.append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
    return float(loss.detach()) if phases is not None else None


for i in range(3):
    step(i)
torch.cuda.synchronize()
ph = []
for i in range(10):
    step(3 + i, ph)
ph = np.array(ph)
res = {"rays_per_step": R, "tetrahedra": int(len(C)), "config": "tetra-nerf (128+128, biased), train mode (stratified jitter), Adam",
       "forward_ms": round(float(np.median(ph[:, 0])), 3), "backward_ms": round(float(np.median(ph[:, 1])), 3), "optimizer_ms": round(float(np.median(ph[:, 2])), 3),
       "step_ms": round(float(np.median(ph.sum(1))), 3), "rays_per_s": round(R / float(np.median(ph.sum(1))) * 1e3),
       "note": "unfused training path: this repo's CUDA API ops + torch fp32 MLP/autograd (baseline for the fused backward of SURVEY 8f-1)"}
print(res)
os.makedirs(R_ + "/gpurun_out", exist_ok=True)
json.dump(res, open(R_ + "/gpurun_out/r1_train_step.json", "w"), indent=1)
