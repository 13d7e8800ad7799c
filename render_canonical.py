#!/usr/bin/env python
"""360-degree orbit render of a canonical avatar -- the reference's render_canonical.py (:37-137) on
avatarcraft_b200.  Same options (subset); `--weights_path` is a reference-layout state dict
(`torch.save(net.state_dict())`, stylize.py:255-260).  Without one, `--synthetic` renders the synthetic
"trained-like" checkpoint (the released avatars are Google-Drive downloads, readme.md:55,72).

    python render_canonical.py --synthetic --exp_name demo --render_h 256 --render_w 256 --n_views 8
"""
import argparse
import os

import torch

from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import render_utils, synthetic
from avatarcraft_b200.utils.camera_paths import default_360_path, rays_for_pose
from avatarcraft_b200.utils.checkpoint import AsyncImageWriter
from avatarcraft_b200.utils.constant import CAN_HEAD_CAMERA_DIST, CAN_HEAD_OFFSET, CANONICAL_CAMERA_DIST_VAL, NSR_BOUND


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights_path", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--exp_name", type=str, default="canonical")
    ap.add_argument("--render_h", type=int, default=256)
    ap.add_argument("--render_w", type=int, default=256)
    ap.add_argument("--n_views", type=int, default=60, help="trajectory_resolution of the orbit (reference: 60)")
    ap.add_argument("--rays_per_batch", type=int, default=4096)
    ap.add_argument("--render_head", action="store_true", help="also render the head close-up orbit")
    opt = ap.parse_args()

    net = NeRFNetwork()
    if opt.weights_path:
        net.load_state_dict(torch.load(opt.weights_path, map_location="cpu"))
    elif opt.synthetic:
        net.load_state_dict(synthetic.synthetic_state_dict("trained", 43))
    else:
        raise SystemExit("give --weights_path or --synthetic")
    net = net.cuda().eval()
    out_dir = os.path.join("demo", "canonical_360", opt.exp_name)
    os.makedirs(out_dir, exist_ok=True)
    orbits = [("body", (0.0, 0.0, 0.0), CANONICAL_CAMERA_DIST_VAL + 0.26)]         # render_canonical.py:47-49 uses 1.7
    if opt.render_head:
        orbits.append(("head", (0.0, CAN_HEAD_OFFSET, 0.0), CAN_HEAD_CAMERA_DIST))
    for tag, center, dist in orbits:
        # rays are generated on the device (ac_gen_rays); PNG/GIF encoding runs on a worker thread behind an async D2H copy
        writer = AsyncImageWriter(gif_path=os.path.join(out_dir, f"{opt.exp_name}_{tag}.gif"))
        poses = default_360_path(center, dist, opt.n_views)
        for i, pose in enumerate(poses):
            o, d = rays_for_pose(pose, opt.render_w, opt.render_h, "cuda")
            with torch.no_grad():
                rgb, _ = render_utils.render_instantnsr_naive(net, o, d, opt.rays_per_batch, requires_grad=False, render_can=True,
                                                              perturb=False, bound=NSR_BOUND)
            writer.submit(rgb.reshape(opt.render_h, opt.render_w, 3), os.path.join(out_dir, f"{opt.exp_name}_{tag}_{i:04d}.png"))
        writer.close()
        print(f"{tag}: {len(poses)} views -> {out_dir}")


if __name__ == "__main__":
    main()
