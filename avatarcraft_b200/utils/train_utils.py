"""One optimisation step of the stylisation trainer, restated from stylize.py:135-199 (pass 2 of the
two-pass SDS scheme): for every `batch_size`-ray patch re-render WITH gradients, back-propagate
  (a) the cached pixel gradient,  rgb_pred_patch.backward(gradient=rgb_global_grad[patch])   (:163)
  (b) w_eikonal * eikonal                                                                      (:166-170)
  (c) 1e5 * smooth_l1(clamp(opacity_pred), clamp(opacity_gt))  against the frozen net_gt        (:177-193)
then optimizer.step().  Multi-GPU: patches are sharded over ranks and the gradients all-reduced once
(utils/distributed.py).  The SDS pixel gradient itself comes from the Stable-Diffusion guidance
(models/diffusion.py:92-149), which is third-party arithmetic; any [n_rays,3] tensor can be supplied."""
import torch
import torch.nn.functional as F

from .constant import NSR_BOUND, WHITE_BKG
from .distributed import allreduce_gradients, masked_mean_share, shard_patches
from .optim import FlatAdam
from .render_utils import render_instantnsr_naive


def stylize_patch_step(net_style, net_gt, optimizer, rays_o, rays_d, pixel_grad, batch_size=4096, w_eikonal=0.01,
                       use_opacity=True, bkg_key=WHITE_BKG, rank=0, world=1, jitter=None):
    """rays_o/rays_d [n,3] (device), pixel_grad [n,3] = d(SDS loss)/d(rgb).  Returns a dict of detached
    scalars (eikonal mean, opacity loss).  Semantics of stylize.py:143-199 with perturb=1.0 for the style
    net (training mode -> jitter) and the eval-mode frozen net_gt."""
    n = rays_o.shape[0]
    batch_size = min(batch_size, n)
    optimizer.zero_grad()
    stats = {"eikonal": [], "opacity": []}
    for s, e, scale in shard_patches(n, batch_size, rank, world):
        o, d = rays_o[s:e], rays_d[s:e]
        rgb, eik, extra = render_instantnsr_naive(net_style, o, d, requires_grad=True, bkg_key=bkg_key, return_torch=True,
                                                  rays_per_batch=batch_size, perturb=1.0, return_raw=True, render_can=True,
                                                  bound=NSR_BOUND)
        loss = (rgb * pixel_grad[s:e]).sum()                       # == rgb.backward(gradient=pixel_grad)
        if w_eikonal > 0.0:
            # a split patch: the eikonal term is a MASKED mean (|p| < 1.2), so each rank's share is its mask count over the
            # patch's, not its ray count (one scalar all-reduce); whole patches keep weight 1
            share = masked_mean_share(net_style.last_eikonal_count) if scale != 1.0 else 1.0
            loss = loss + eik * (w_eikonal * share)
            stats["eikonal"].append(eik.detach() * share)
        if use_opacity and net_gt is not None:
            with torch.no_grad():
                _, _, extra_gt = render_instantnsr_naive(net_gt, o, d, requires_grad=False, bkg_key=bkg_key, return_torch=True,
                                                         rays_per_batch=batch_size, perturb=True, return_raw=True,
                                                         render_can=True)
            op = F.smooth_l1_loss(extra["weight_sum"].clamp(0.0, 1.0), extra_gt["weight_sum"].clamp(0.0, 1.0)) * 1e5
            loss = loss + op * scale
            stats["opacity"].append(op.detach())
        loss.backward()
    if isinstance(optimizer, FlatAdam):            # flat buffers: ONE in-place all-reduce + ONE update launch
        optimizer.all_reduce()
        optimizer.step()
    else:
        allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]])
        optimizer.step()
    return {k: (torch.stack(v).mean() if v else None) for k, v in stats.items()}
