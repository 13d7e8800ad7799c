"""Ray generation for the warp entry point: the NeuS-convention rays of
utils/SMPLDataset.py:86-103 `gen_rays_pose` -- p = ((x-cx)/f, -(y-cy)/f, -1), normalised, rotated by the
camera-to-world pose; pixel grid linspace(0, W-1, W//l)."""
import numpy as np
import torch


def dataset_intrinsics(H=512, W=512, camera_angle_x=1.0472):
    """Intrinsics SMPLDataset builds for data/smpl_da_512 (focal = 0.5*W / tan(0.5*camera_angle_x))."""
    f = 0.5 * W / np.tan(0.5 * camera_angle_x)
    return torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=torch.float32)


def gen_rays_pose(pose, K, H, W, resolution_level=1, device="cpu"):
    """pose [4,4] camera-to-world -> rays_o, rays_v [H//l, W//l, 3]."""
    l = resolution_level
    pose = torch.as_tensor(pose, dtype=torch.float32, device=device)
    K = K.to(device)
    tx = torch.linspace(0, W - 1, int(W // l), device=device)
    ty = torch.linspace(0, H - 1, int(H // l), device=device)
    px, py = torch.meshgrid(tx, ty, indexing="ij")
    px, py = px.t(), py.t()
    p = torch.stack([(px - K[0][2]) / K[0][0], -(py - K[1][2]) / K[1][1], -torch.ones_like(px)], -1).float()
    v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    v = torch.sum(v[..., None, :] * pose[:3, :3], -1)
    return pose[None, None, :3, 3].expand(v.shape), v
