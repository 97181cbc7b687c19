"""Ray-sharded rendering over the GPUs of one node (one process per GPU, torch.distributed).

The path shards naturally: rays are independent, the read-only mesh / field / MLP are replicated
on every rank, each rank renders a contiguous slice of the ray batch and the rendered pixels (rgb 3, accumulation 1,
depth 1, mask 1 -> 6 floats per ray) are exchanged with ONE all-gather (NCCL over NVLink on GPUs, gloo in the CPU
tests).  There is no collective on the data path itself.  The reference's only multi-GPU mechanism is nerfstudio's
DDP wrap for training (tetranerf/nerfstudio/pipeline.py:53-58)."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced: the first (num_rays % world) ranks get one extra ray"""
    base, rem = divmod(num_rays, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_render(render_fn: Callable[[torch.Tensor, torch.Tensor], Dict[str, torch.Tensor]], origins: torch.Tensor,
                   directions: torch.Tensor, group=None) -> Dict[str, torch.Tensor]:
    """Every rank passes the FULL ray batch (or at least its own slice filled in); returns the full-batch outputs on
    every rank.  `render_fn(o, d)` renders a slice on the local device and returns rgb[R,3], accumulation[R,1],
    depth[R,1], ray_mask[R]."""
    if not dist.is_available() or not dist.is_initialized():
        return render_fn(origins, directions)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    R = origins.shape[0]
    lo, hi = shard_bounds(R, rank, world)
    part = render_fn(origins[lo:hi].contiguous(), directions[lo:hi].contiguous())
    per = (R + world - 1) // world  # equal-size buffers for all_gather_into_tensor
    pack = torch.zeros((per, 6), dtype=torch.float32, device=origins.device)
    n = hi - lo
    if n:
        pack[:n, 0:3] = part["rgb"]
        pack[:n, 3:4] = part["accumulation"]
        pack[:n, 4:5] = part["depth"]
        pack[:n, 5] = part["ray_mask"].to(torch.float32)
    gathered = torch.empty((world * per, 6), dtype=torch.float32, device=origins.device)
    dist.all_gather_into_tensor(gathered, pack, group=group)
    rows = torch.cat([torch.arange(r * per, r * per + (shard_bounds(R, r, world)[1] - shard_bounds(R, r, world)[0]), device=origins.device)
                      for r in range(world)])
    full = gathered[rows]
    return {"rgb": full[:, 0:3].contiguous(), "accumulation": full[:, 3:4].contiguous(), "depth": full[:, 4:5].contiguous(),
            "ray_mask": full[:, 5] > 0.5}


def average_gradients(parameters, group=None) -> None:
    """DDP semantics of the reference's only multi-GPU mechanism (tetranerf/nerfstudio/pipeline.py:53-58 wraps the model in
    DistributedDataParallel): every rank draws its own ray batch, the gradients are AVERAGED over the ranks before the optimizer step --
    `tetrahedra_field.grad` ([64,V]: 77 MB at 300k vertices) and the twelve MLP gradients, packed into one flat buffer and reduced with a
    single all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The fused training step produces all gradients at the end of
    one backward op, so there is nothing to bucket or overlap inside it."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
