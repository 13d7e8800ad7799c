"""Ray-batch render driver with the reference's signature
(utils/render_utils.py:514-600 render_instantnsr_naive, :953-987 select_background) plus the
camera-orbit helpers the entry points need (:57-76 pose_spherical-style, :137-154).

B200-first difference: the reference loops over `rays_per_batch` chunks, launching ~250 eager
kernels per chunk and concatenating.  Here all rays of the call go through ONE fused launch;
`rays_per_batch` only defines the segments over which the eikonal mean is taken, so the returned
`total_eikonal` (sum of per-batch means, :575) is unchanged."""
import numpy as np
import torch

from .constant import WHITE_BKG, BLACK_BKG, NOISE_BKG, CHESSBOARD_BKG


def select_background(shape, key) -> torch.Tensor:
    """Background colours for `shape=(n_rays, 3)`: white / black / per-ray gaussian grey
    (mean 0.5, std 0.1, clamped) / blurred chessboard (utils/render_utils.py:953-987)."""
    key = key % 4
    n = shape[0]
    if key == WHITE_BKG:
        return torch.ones(shape)
    if key == BLACK_BKG:
        return torch.zeros(shape)
    if key == NOISE_BKG:
        grey = torch.nn.init.normal_(torch.ones(n), mean=0.5, std=0.1).clamp_(0, 1)
        return grey[:, None].expand(n, 3).contiguous()
    side = int(np.sqrt(n))                       # chessboard assumes a square batch
    cell = max(side // 10, 1)
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing='xy')
    board = np.where(((ii // cell) + (jj // cell)) % 2 == 0, 0.8, 0.2).astype(np.float32).T
    from torchvision import transforms           # optional dependency, only for this background
    blur = transforms.GaussianBlur(kernel_size=(5, 9), sigma=(0.1, 2.0))
    img = blur(torch.from_numpy(board)[None, None])[0, 0]
    return img.reshape(-1, 1).expand(side * side, 3).contiguous()


def select_background_device(n_rays, key, device, seed=None, sigma=None) -> torch.Tensor:
    """select_background on the device (one kernel, no host tensor, no H2D): [n_rays,3].  The noise background uses a
    counter-based generator keyed by `seed` (drawn from torch's generator when None); the chessboard blur width
    `sigma` is drawn U(0.1, 2.0) per call like torchvision's GaussianBlur when None."""
    from .. import _lib
    key = key % 4
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if key == NOISE_BKG else 0
    if sigma is None:
        sigma = float(torch.empty(1).uniform_(0.1, 2.0).item()) if key == CHESSBOARD_BKG else 1.0
    out = torch.empty(n_rays, 3, device=device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().ac_select_background(int(key), int(n_rays), int(seed), float(sigma), _lib.ptr(out), _lib.stream_ptr()),
                   "ac_select_background")
    return out


def render_instantnsr_naive(net, rays_o, rays_d, rays_per_batch=6400, requires_grad=False, return_torch=True,
                            bkg_key: int = WHITE_BKG, render_can: bool = False, perturb: bool = True,
                            return_raw: bool = False, verts=None, faces=None, Ts=None, num_steps: int = 64,
                            upsample_steps=64, bound: float = 1.6):
    """rays_o, rays_d [H*W,3] -> rgb [H*W,3], total_eikonal (and extra_out with depth [.,1],
    weight_sum [.,1], normal [.,3] when return_raw) -- utils/render_utils.py:514-600."""
    device = rays_o.device
    total = rays_o.shape[0]
    key = bkg_key % 4
    if key == WHITE_BKG:
        bg = None                                                   # kernel default: white
    elif key == BLACK_BKG:
        bg = torch.zeros(total, 3, device=device)
    else:                                                           # generated per batch, like the reference
        bg = torch.cat([select_background_device(min(rays_per_batch, total - i), bkg_key, device)
                        for i in range(0, total, rays_per_batch)])
    with torch.set_grad_enabled(requires_grad):
        if requires_grad and total > rays_per_batch:
            raise NotImplementedError("requires_grad=True renders one patch per call (as stylize.py:153-158 does)")
        out = net.render(rays_o.unsqueeze(0), rays_d.unsqueeze(0), num_steps=num_steps, upsample_steps=upsample_steps,
                         bound=bound, staged=False, bg_color=bg, cos_anneal_ratio=1.0, normal_epsilon_ratio=0.0,
                         render_can=render_can, verts=verts, faces=faces, Ts=Ts, perturb=perturb,
                         per_sample_outputs=False, eikonal_segment=rays_per_batch)
        rgb = out['rgb'].reshape(-1, 3)
        extra_out = {"depth": out['depth'].reshape(-1, 1), "weight_sum": out['weight_sum'].reshape(-1, 1),
                     "normal": out['normal'].reshape(-1, 3)}
        total_eikonal = out['gradient_error'].sum()
    if not return_torch:
        rgb = rgb.detach().cpu().numpy()
        extra_out = {k: v.detach().cpu().numpy() for k, v in extra_out.items()}
    if return_raw:
        return rgb, total_eikonal, extra_out
    return rgb, total_eikonal
