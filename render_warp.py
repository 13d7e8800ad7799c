#!/usr/bin/env python
"""Animate / reshape a canonical avatar through the SMPL-guided inverse warp -- the reference's render_warp.py
(:25-124) on avatarcraft_b200: per frame the host computes the posed SMPL surface and per-vertex transforms
(models/smpl.py::calc_local_trans), every per-sample step runs on the GPU (no libigl / numpy round trips).

The SMPL model pickle, smpl_uv.obj and AMASS clips are not redistributable (readme.md:41-59); `--synthetic`
substitutes the SMPL-shaped synthetic body and a sinusoidal pose clip.

    python render_warp.py --synthetic --exp_name demo --resolution 128 --max_frames 4
"""
import argparse
import os

import numpy as np
import torch

from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.models.smpl import SMPL, calc_local_trans
from avatarcraft_b200.utils import render_utils, synthetic
from avatarcraft_b200.utils.constant import BLACK_BKG, NSR_BOUND, WHITE_BKG
from avatarcraft_b200.utils.ray_gen import dataset_intrinsics, gen_rays_pose


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights_path", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--smpl_model", type=str, default="data/smplx/smpl/SMPL_NEUTRAL.pkl")
    ap.add_argument("--poseseq_path", type=str, default=None)
    ap.add_argument("--render_type", type=str, default="animate", choices=["animate", "interp_shape"])
    ap.add_argument("--exp_name", type=str, default="warp")
    ap.add_argument("--resolution", type=int, default=256)
    ap.add_argument("--max_frames", type=int, default=120)
    ap.add_argument("--white_bkg", type=int, default=1)
    opt = ap.parse_args()

    net = NeRFNetwork()
    if opt.synthetic:
        net.load_state_dict(synthetic.synthetic_state_dict("trained", 43))
        body = SMPL(synthetic.synthetic_smpl_model())
        poses = synthetic.sinusoid_pose_sequence(opt.max_frames)
        faces = synthetic.synthetic_body()["faces"]
    else:
        net.load_state_dict(torch.load(opt.weights_path, map_location="cpu"))
        body = SMPL(opt.smpl_model)
        poses = np.load(opt.poseseq_path).astype(np.float32).reshape(-1, 72) if opt.poseseq_path else None
        faces = np.concatenate([body.faces, body.faces], 1)
    net = net.cuda().eval()
    net.warp_skip_masked = True          # only img is kept below (as render_warp.py:88): masked-out samples need no search / evaluation
    shape_from, shape_to = np.zeros((1, 10), np.float32), np.zeros((1, 10), np.float32)
    shape_from[0, 1], shape_to[0, 1] = 2.0, -2.0                      # render_warp.py:37-45 defaults
    world_verts, Ts, n_frames = calc_local_trans(body, poses=poses, shape_from=shape_from, shape_to=shape_to,
                                                 render_type=opt.render_type, max_frames=opt.max_frames)
    cam = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 2.4], [0, 0, 0, 1]], np.float32)     # front view, NeuS convention (-z forward)
    K = dataset_intrinsics()
    out_dir = os.path.join("demo", "test_views", opt.exp_name)
    os.makedirs(out_dir, exist_ok=True)
    from PIL import Image
    frames = []
    for i in range(n_frames):
        o, d = gen_rays_pose(cam, K, 512, 512, resolution_level=512 // opt.resolution, device="cuda")
        o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
        with torch.no_grad():
            img, _, _ = render_utils.render_instantnsr_naive(net, o, d, 64 * 128, requires_grad=False,
                                                             bkg_key=WHITE_BKG if opt.white_bkg else BLACK_BKG, perturb=False,
                                                             return_raw=True, render_can=False, verts=world_verts[i], faces=faces,
                                                             Ts=Ts[i], num_steps=32, upsample_steps=32, bound=NSR_BOUND)
        im = Image.fromarray((img.reshape(opt.resolution, opt.resolution, 3).clamp(0, 1).cpu().numpy() * 255 + 0.5).astype(np.uint8))
        im.save(os.path.join(out_dir, f"{opt.exp_name}_{i:04d}.png"))
        frames.append(im)
    frames[0].save(os.path.join(out_dir, f"{opt.exp_name}.gif"), save_all=True, append_images=frames[1:], duration=100, loop=0)
    print(f"{n_frames} frames -> {out_dir}")


if __name__ == "__main__":
    main()
