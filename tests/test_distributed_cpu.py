"""world_size-2 gloo test (CPU) of the ray-shard + pixel all-gather plumbing used for N > 1 GPUs."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _fake_render(o, d):
    """deterministic stand-in for the per-rank CUDA render: any function of the rays alone"""
    rgb = torch.sigmoid(o * 3 + d)
    acc = (o.sum(-1, keepdim=True) * 0.1).sin().abs()
    depth = (o - d).norm(dim=-1, keepdim=True)
    return {"rgb": rgb, "accumulation": acc, "depth": depth, "ray_mask": d[:, 0] > 0}


def _worker(rank, world, port, R, out_dir):
    for p in (str(ROOT), str(ROOT / "tetra-nerf_b200")):
        sys.path.insert(0, p)
    from tetranerf.b200.distributed import shard_bounds, sharded_render

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    o, d = torch.randn((R, 3), generator=g), torch.randn((R, 3), generator=g)
    out = sharded_render(_fake_render, o, d)
    ref = _fake_render(o, d)
    ok = all(torch.equal(out[k], ref[k]) for k in ref)
    lo, hi = shard_bounds(R, rank, world)
    torch.save({"ok": ok, "lo": lo, "hi": hi}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("R", [10, 7, 1])
def test_sharded_render_gloo_world2(tmp_path, R):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), R, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == R  # contiguous, covering


def test_shard_bounds_properties():
    sys.path.insert(0, str(ROOT / "tetra-nerf_b200"))
    from tetranerf.b200.distributed import shard_bounds

    for R in (0, 1, 5, 4096, 65537):
        for W in (1, 2, 3, 8):
            b = [shard_bounds(R, r, W) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == R and all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


# ---- DDP gradient semantics of the training step (tetranerf/nerfstudio/pipeline.py:53-58) ----------------------------------------
def _grad_worker(rank, world, port, out_dir):
    for p in (str(ROOT), str(ROOT / "tetra-nerf_b200")):
        sys.path.insert(0, p)
    import numpy as np

    from oracle import oracle as orc
    from tetranerf.b200 import synthetic as syn
    from tetranerf.b200.distributed import average_gradients, shard_bounds

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    V, C = syn.delaunay_mesh(800, seed=2)
    mesh = orc.OracleMesh(V, C)
    cfg = orc.RenderConfig(num_samples=16, num_fine_samples=16, use_biased_sampler=True)
    R = 32
    o, d = syn.camera_rays(R, seed=3)
    g = torch.Generator().manual_seed(4)
    jc, jf, tgt = torch.rand((R, 17), generator=g), torch.rand((R, 17), generator=g), torch.rand((R, 3), generator=g)

    def grads(lo, hi):
        f = torch.from_numpy(syn.random_field(len(V), 64, seed=3)).requires_grad_(True)
        p = {k: v.clone().requires_grad_(True) for k, v in orc.init_mlp_params(0).items()}
        out = orc.render_train(mesh, f, p, o[lo:hi], d[lo:hi], cfg, jc[lo:hi], jf[lo:hi], nthreads=1)
        torch.nn.functional.mse_loss(out["rgb"], tgt[lo:hi]).backward()
        return [f] + [p[k] for k in sorted(p)]

    lo, hi = shard_bounds(R, rank, world)
    mine = grads(lo, hi)            # this rank's own batch (equal sizes: the mean of the per-rank mean losses = the loss of the whole batch)
    average_gradients(mine)
    whole = grads(0, R)
    err = max(((a.grad - b.grad).abs().max() / b.grad.abs().max().clamp_min(1e-30)).item() for a, b in zip(mine, whole))
    torch.save({"err": err}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_ddp_gradient_average_equals_whole_batch_gradient(tmp_path):
    """two ranks, each differentiating its half of the batch through the training-mode oracle render; after average_gradients every
    rank holds the gradient of the concatenated batch (field and all twelve MLP parameters)"""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    errs = [torch.load(tmp_path / f"g{r}.pt")["err"] for r in range(world)]
    assert max(errs) < 1e-5, errs
