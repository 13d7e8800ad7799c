#!/usr/bin/env python
"""Stylisation trainer -- the optimisation loop of the reference's stylize.py (Trainer.train :47-217) on
avatarcraft_b200: per view, pass 1 renders the (sub-sampled) image without gradients, a guidance function turns it
into a pixel gradient, pass 2 re-renders 4096-ray patches with gradients and back-propagates pixel gradient +
eikonal + opacity-vs-frozen-copy, ONE gradient all-reduce (multi-GPU), Adam(lr 5e-3).

`--guidance` selects what turns the pass-1 image into a pixel gradient:
  sds      score distillation through avatarcraft_b200.models.diffusion.StableDiffusion (models/diffusion.py:92-149):
           VAE encode with gradient, UNet on the native tcgen05 kernels, classifier-free guidance.  The HF weights are
           not available offline: `--sd_weights <diffusers dir>` loads them, otherwise the networks are random-init
           (right cost, meaningless images)
  target   pulls the render towards a fixed colour tint (deterministic, for smoke tests)
  randn    unit Gaussian pixel gradient
Launch with torchrun for multi-GPU: patches are sharded across ranks.

    python stylize.py --synthetic --exp_name demo --n_views 4 --coarse_epochs 1 --fine_epochs 0
"""
import argparse
import os

import torch
import torch.distributed as dist

from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import render_utils, synthetic
from avatarcraft_b200.utils.camera_paths import default_360_path, rays_for_pose
from avatarcraft_b200.utils.constant import CANONICAL_CAMERA_DIST_TRAIN, NSR_BOUND
from avatarcraft_b200.utils.checkpoint import load_checkpoint, save_checkpoint
from avatarcraft_b200.utils.distributed import render_rays_sharded
from avatarcraft_b200.utils.optim import FlatAdam
from avatarcraft_b200.utils.train_utils import stylize_patch_step


def guidance(kind, rgb_chw, step):
    if kind == "randn":
        return torch.randn(rgb_chw.shape, device=rgb_chw.device, generator=torch.Generator(rgb_chw.device).manual_seed(step))
    tint = torch.tensor([0.9, 0.6, 0.3], device=rgb_chw.device).view(1, 3, 1, 1)
    return (rgb_chw - tint) / rgb_chw[0, 0].numel()                    # d/d(rgb) of 0.5 * mean squared error to the tint


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights_path", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--exp_name", type=str, default="style")
    ap.add_argument("--render_h", type=int, default=256)
    ap.add_argument("--render_w", type=int, default=256)
    ap.add_argument("--n_views", type=int, default=100)
    ap.add_argument("--coarse_epochs", type=int, default=40)
    ap.add_argument("--fine_epochs", type=int, default=20)
    ap.add_argument("--subsample_scale", type=int, default=4)
    ap.add_argument("--batch_size", type=int, default=4096)
    ap.add_argument("--w_eikonal", type=float, default=0.01)
    ap.add_argument("--use_opacity", type=int, default=1)
    ap.add_argument("--guidance", type=str, default="target", choices=["target", "randn", "sds"])
    ap.add_argument("--prompt", type=str, default="a 3D rendering of a knight in bronze armour")
    ap.add_argument("--sd_version", type=str, default="1.5", choices=["1.5", "2.0"])
    ap.add_argument("--sd_weights", type=str, default=None, help="diffusers-format directory (unet/, vae/); random init when absent")
    ap.add_argument("--guidance_scale", type=float, default=100.0)
    ap.add_argument("--i_save", type=int, default=1000)
    ap.add_argument("--lr", type=float, default=5e-3)
    ap.add_argument("--lr_decay", action="store_true",
                    help="halve the LR after half of the epochs.  Off by default: the reference builds a StepLR but never steps it "
                         "(`# scheduler.step()`, stylize.py:214), so its LR stays at 5e-3 for the whole run")
    ap.add_argument("--resume", type=str, default=None, help="a *.pth.tar written by this script (its .resume.pt is picked up)")
    opt = ap.parse_args()

    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sd = synthetic.synthetic_state_dict("trained", 43) if opt.synthetic else torch.load(opt.weights_path, map_location="cpu")
    net_style, net_gt = NeRFNetwork(), NeRFNetwork()
    net_style.load_state_dict(sd); net_gt.load_state_dict(sd)
    net_style, net_gt = net_style.cuda().train(), net_gt.cuda().eval()        # net_style is never .eval()ed (stylize.py:336)
    for p in net_gt.parameters():
        p.requires_grad_(False)
    # Adam(lr 5e-3) (stylize.py:355-363; its StepLR is never stepped, :214) on flat buffers: one all-reduce + one update launch
    optimizer = FlatAdam(net_style.parameters(), lr=opt.lr)
    sd_guide = text_emb = None
    if opt.guidance == "sds":                                                  # stylize.py:340-352 setup_loss + :100 text embeds
        from avatarcraft_b200.models.diffusion import StableDiffusion
        sd_guide = StableDiffusion("cuda", opt.sd_version, weights_dir=opt.sd_weights)
        sd_guide.cfg_parallel = world > 1       # every rank draws (t, noise) from the per-step seed below, so the pair may be split
        text_emb = sd_guide.get_text_embeds(opt.prompt)
    out_dir = os.path.join("style", "canonical_360", opt.exp_name)
    os.makedirs(out_dir, exist_ok=True)
    H, W, step, first_epoch, first_view = opt.render_h, opt.render_w, 0, 0, 0
    n_epochs = opt.coarse_epochs + opt.fine_epochs
    if opt.resume:                                                             # weights + optimizer moments + counters + RNG
        info = load_checkpoint(opt.resume, net_style, optimizer)
        step, first_epoch = info["step"], info["epoch"]
        first_view = int(info["extra"].get("view_pos", 0))                     # position inside the epoch's view permutation
        print(f"resumed from {opt.resume}: step {step}, epoch {first_epoch}, view {first_view}, "
              f"optimizer state {'restored' if info['resumed'] else 'fresh'}")
        if first_epoch >= n_epochs:
            print("the checkpoint is from a finished run: nothing left to train")
    for epoch in range(first_epoch, n_epochs):
        lr = opt.lr * (0.5 ** (epoch // max(n_epochs // 2, 1))) if opt.lr_decay else opt.lr
        optimizer.param_groups[0]["lr"] = lr
        stride = opt.subsample_scale if epoch < opt.coarse_epochs else min(1, opt.subsample_scale // 2) or 1
        poses = default_360_path((0.0, 0.0, 0.0), CANONICAL_CAMERA_DIST_TRAIN, opt.n_views)
        perm = torch.randperm(opt.n_views, generator=torch.Generator().manual_seed(epoch)).tolist()      # same order on every rank
        start = first_view if epoch == first_epoch else 0                      # a resumed epoch continues where it stopped
        for vpos in range(start, len(perm)):
            vi = perm[vpos]
            o, d = rays_for_pose(poses[vi], W, H, "cuda")
            o, d = o.reshape(H, W, 3)[::stride, ::stride].reshape(-1, 3).contiguous(), d.reshape(H, W, 3)[::stride, ::stride].reshape(-1, 3).contiguous()
            h, w = H // stride, W // stride
            with torch.no_grad():                                                                # pass 1 (stylize.py:115), ray-sharded
                fn = lambda a, b: render_utils.render_instantnsr_naive(net_style, a, b, opt.batch_size, render_can=True, perturb=True)[0]
                rgb = render_rays_sharded(fn, o, d, rank, world)                                 # pads a ray count that does not divide
            if sd_guide is not None:                                                             # SDS (models/diffusion.py:92-149)
                pixel_grad = sd_guide.pixel_gradient(text_emb, rgb, h, w, opt.guidance_scale, seed=1_000_003 * (epoch + 1) + step)
            else:
                g = guidance(opt.guidance, rgb.reshape(h, w, 3).permute(2, 0, 1)[None], step)    # [1,3,h,w]
                pixel_grad = g[0].permute(1, 2, 0).reshape(-1, 3).contiguous()
            stats = stylize_patch_step(net_style, net_gt, optimizer, o, d, pixel_grad, batch_size=opt.batch_size,
                                       w_eikonal=opt.w_eikonal, use_opacity=bool(opt.use_opacity), rank=rank, world=world)
            step += 1
            if rank == 0 and step % 10 == 0:
                print(f"epoch {epoch} step {step} eikonal {float(stats['eikonal'] or 0):.4f}")
            if rank == 0 and step % opt.i_save == 0:                           # reference name: <exp>_<step+1 (0-based), 4 digits> (stylize.py:206,256)
                nxt = (epoch, vpos + 1) if vpos + 1 < len(perm) else (epoch + 1, 0)
                save_checkpoint(os.path.join(out_dir, f"{opt.exp_name}_{step:04d}.pth.tar"), net_style, optimizer, step, nxt[0],
                                extra={"view_pos": nxt[1]})
    if rank == 0:
        # the reference's final log_model(global_step) (stylize.py:216) + a fixed name for the render scripts
        save_checkpoint(os.path.join(out_dir, f"{opt.exp_name}_{step + 1:04d}.pth.tar"), net_style, optimizer, step, n_epochs, extra={"view_pos": 0})
        save_checkpoint(os.path.join(out_dir, f"{opt.exp_name}.pth.tar"), net_style, optimizer, step, n_epochs, extra={"view_pos": 0})
        print("saved", os.path.join(out_dir, f"{opt.exp_name}.pth.tar"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
