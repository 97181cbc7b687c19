"""GPU: the drop-in TetrahedraNerf.get_outputs (fused in eval, unfused autograd path in training) vs the oracle render."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tetranerf.b200 import synthetic as syn
from tetranerf.nerfstudio import model as M

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def build_model(V, C, field, **cfg):
    config = M.TetrahedraNerfConfig(num_tetrahedra_vertices=len(V), num_tetrahedra_cells=len(C), **cfg)
    m = M.TetrahedraNerf(config)
    params = orc.init_mlp_params(0)
    sd = {"tetrahedra_vertices": torch.from_numpy(V), "tetrahedra_cells": torch.from_numpy(C), "tetrahedra_field": torch.from_numpy(field)}
    sd.update(params)
    missing = m.load_state_dict(sd, strict=False)
    assert not [k for k in missing.missing_keys if not k.startswith("device_indicator")], missing
    return m.to(DEV), params


@pytest.mark.parametrize("biased", [True, False])
def test_get_outputs_eval_and_train_paths(small_mesh, biased):
    V, C = small_mesh
    field = syn.random_field(len(V), 64, seed=3)
    ns = 128 if biased else 64
    m, params = build_model(V, C, field, num_samples=ns, num_fine_samples=ns, use_biased_sampler=biased)
    o, d = syn.camera_rays(200)
    o[3] = [5, 5, 5]; d[3] = [1, 0, 0]
    ocfg = orc.RenderConfig(num_samples=ns, num_fine_samples=ns, use_biased_sampler=biased)
    ref = orc.render(orc.OracleMesh(V, C), torch.from_numpy(field), params, o, d, ocfg)
    bundle = M.RayBundle(origins=torch.from_numpy(o).to(DEV), directions=torch.from_numpy(d).to(DEV))
    # eval + no_grad -> fused CUDA pipeline
    m.eval()
    with torch.no_grad():
        out = m(bundle)
    assert m._fused is not None
    assert set(out) >= {"rgb", "accumulation", "depth", "ray_mask"}
    assert torch.equal(out["ray_mask"].cpu(), ref["ray_mask"])
    assert (out["rgb"].cpu() - ref["rgb"]).abs().max() < 1e-4 and (out["accumulation"].cpu() - ref["accumulation"]).abs().max() < 1e-4
    # eval with autograd enabled -> the reference's op sequence on our CUDA ops (torch fp32 MLP): also within tolerance
    out_u = m(M.RayBundle(origins=torch.from_numpy(o).to(DEV), directions=torch.from_numpy(d).to(DEV)))
    assert (out_u["rgb"].detach().cpu() - ref["rgb"]).abs().max() < 1e-4
    assert (out_u["rgb"].detach() - out["rgb"]).abs().max() < 1e-4
    # gradients reach the vertex field and the MLP (extension/__init__.py autograd wrapper)
    loss = m.get_loss_dict(out_u, {"image": torch.zeros((len(o), 3))})["rgb_loss"]
    loss.backward()
    g = m.tetrahedra_field.grad
    assert g is not None and g.shape == (64, len(V)) and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert m.mlp_base.layers[0].weight.grad is not None and float(m.mlp_head.layers[0].weight.grad.abs().sum()) > 0
    # a parameter update invalidates the fused renderer's packed copies
    with torch.no_grad():
        m.tetrahedra_field.add_(0.01)
        out2 = m(bundle)
    assert (out2["rgb"] - out["rgb"]).abs().max() > 0
    # training mode (stratified jitter) runs and stays finite
    m.train()
    out_t = m(M.RayBundle(origins=torch.from_numpy(o).to(DEV), directions=torch.from_numpy(d).to(DEV)))
    assert torch.isfinite(out_t["rgb"]).all()
