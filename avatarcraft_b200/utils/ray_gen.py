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


# ---- on-device ray generation: one kernel per view, no host arrays, no H2D ----------------
def _launch_gen_rays(c2w, fx, fy, cx, cy, W, H, x0, xs, y0, ys, convention, device):
    import ctypes
    from .. import _lib
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("device ray generation needs a CUDA device (no CPU path)")
    m = np.ascontiguousarray(np.asarray(c2w, dtype=np.float64).reshape(-1))
    if m.size == 12:
        m = np.concatenate([m, [0.0, 0.0, 0.0, 1.0]])
    if m.size != 16:
        raise RuntimeError("c2w must be a 3x4 or 4x4 camera-to-world matrix")
    with torch.cuda.device(dev):
        o = torch.empty(H * W, 3, device=dev, dtype=torch.float32)
        d = torch.empty(H * W, 3, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().ac_gen_rays(m.ctypes.data_as(ctypes.c_void_p), float(fx), float(fy), float(cx), float(cy), int(W), int(H),
                                          float(x0), float(xs), float(y0), float(ys), int(convention), _lib.ptr(o), _lib.ptr(d),
                                          _lib.stream_ptr()), "ac_gen_rays")
    return o, d


def pinhole_rays_device(c2w, width, height, device="cuda", fx=None, fy=None, cx=None, cy=None):
    """cap2rays / shot_rays (utils/render_utils.py:363-376, utils/ray_utils.py:25-37) on the device: row-major
    integer-pixel rays of a pinhole camera; defaults are render_canonical.py's intrinsics (f = 0.78125 W, c = W/2)."""
    fx = 0.78125 * width if fx is None else fx
    fy = 0.78125 * height if fy is None else fy
    cx = width / 2 if cx is None else cx
    cy = height / 2 if cy is None else cy
    return _launch_gen_rays(c2w, fx, fy, cx, cy, width, height, 0.0, 1.0, 0.0, 1.0, 0, device)


def gen_rays_pose_device(pose, K, H, W, resolution_level=1, device="cuda"):
    """SMPLDataset.gen_rays_pose (utils/SMPLDataset.py:86-103) on the device -> rays_o, rays_v [H//l, W//l, 3]."""
    l = resolution_level
    w, h = int(W // l), int(H // l)
    K = np.asarray(torch.as_tensor(K).cpu(), dtype=np.float64)
    pose = np.asarray(torch.as_tensor(pose).cpu(), dtype=np.float64)
    xs = (W - 1) / (w - 1) if w > 1 else 0.0
    ys = (H - 1) / (h - 1) if h > 1 else 0.0
    o, v = _launch_gen_rays(pose, K[0, 0], K[1, 1], K[0, 2], K[1, 2], w, h, 0.0, xs, 0.0, ys, 1, device)
    return o.reshape(h, w, 3), v.reshape(h, w, 3)
