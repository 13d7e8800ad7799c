"""Ray/patch sharding and the single gradient all-reduce of a multi-GPU step.

The reference is single-GPU (no distributed code anywhere).  Rays are independent, and its trainer
already accumulates gradients over independent `batch_size`-ray patches before one
`optimizer.step()` (stylize.py:143-199), with every regulariser a per-patch mean
(models/instant_nsr.py:270-272, stylize.py:190).  So a data-parallel step is: assign patches
round-robin to ranks, back-propagate locally, ONE all-reduce(SUM) of the flat fp32 gradient
(12 248 902 floats = 49.0 MB), identical Adam update on every rank (no parameter broadcast)."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_patches(n_rays: int, batch_size: int, rank: int, world: int) -> List[Tuple[int, int, float]]:
    """Patches [(start, end, mean_scale)] this rank renders.  Patch-granular round-robin when there are
    at least `world` patches; otherwise the patches are split evenly across ranks and mean-type losses of a
    partial patch must be scaled by mean_scale = local_rays / patch_rays so that the SUM over ranks
    reproduces the reference's per-patch mean."""
    patches = [(s, min(s + batch_size, n_rays)) for s in range(0, n_rays, batch_size)]
    if len(patches) >= world:
        return [(s, e, 1.0) for i, (s, e) in enumerate(patches) if i % world == rank]
    out = []
    for s, e in patches:
        n = e - s
        lo = s + (n * rank) // world
        hi = s + (n * (rank + 1)) // world
        if hi > lo:
            out.append((lo, hi, (hi - lo) / n))
    return out


def shard_rays(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rays for inference (each rank renders its block; 32 B/ray gathered to rank 0)."""
    return (n_rays * rank) // world, (n_rays * (rank + 1)) // world


def allreduce_gradients(params, group=None) -> int:
    """One all-reduce(SUM) over all parameter gradients, flattened into a single fp32 buffer.
    Returns the number of floats reduced.  Parameters without a gradient contribute zeros (their slot is
    still reduced so every rank issues an identical collective)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    dev = params[0].device
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    at = 0
    for p in params:
        n = p.numel()
        g = flat[at:at + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        at += n
    assert flat.device == dev
    return flat.numel()


def render_rays_sharded(render_fn, rays_o, rays_d, rank: int, world: int, group=None):
    """Pass 1 of a multi-GPU step: every rank renders its contiguous block of rays with `render_fn(o, d) -> rgb [m,3]`
    and the blocks are exchanged with ONE all-gather (12 B per ray), so each rank holds the whole image for the
    guidance.  Rays are independent (no cross-ray term in `run`), so the result equals the single-process render."""
    n = rays_o.shape[0]
    if world <= 1 or not (dist.is_available() and dist.is_initialized()):
        return render_fn(rays_o, rays_d)
    pad = (-n) % world                  # a ray count that does not divide: repeat the last ray, drop the copies after the gather
    if pad:
        rays_o = torch.cat([rays_o, rays_o[-1:].expand(pad, -1)])
        rays_d = torch.cat([rays_d, rays_d[-1:].expand(pad, -1)])
    s, e = shard_rays(n + pad, rank, world)
    mine = render_fn(rays_o[s:e].contiguous(), rays_d[s:e].contiguous()).contiguous()
    full = torch.empty(n + pad, mine.shape[1], device=mine.device, dtype=mine.dtype)
    dist.all_gather_into_tensor(full, mine, group=group)
    return full[:n]


def masked_mean_share(local_count: torch.Tensor, group=None) -> torch.Tensor:
    """Weight that turns this rank's masked mean (sum_local / (count_local + 1e-5), the eikonal term of
    models/instant_nsr.py:266-272) into its share of the whole patch's masked mean when ONE patch is split over ranks:
    (count_local + 1e-5) / (count_global + 1e-5), with count_global from one scalar all-reduce.  The sum over ranks of
    share * local_mean is sum_global / (count_global + 1e-5) -- what a single process computes.  (Scaling by
    local_rays / patch_rays is only right for plain means: the mask count differs per rank.)"""
    total = local_count.detach().clone().float()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return (local_count.detach().float() + 1e-5) / (total + 1e-5)
