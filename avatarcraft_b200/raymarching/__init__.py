"""Occupancy-grid ray marching operators with the reference's names (`raymarching/raymarching.py:21-188`):
march_rays_train, composite_rays_train (differentiable), march_rays, composite_rays, compact_rays.
The reference never imports its own module (dead code); these exist for API completeness
and are validated on the GPU box against the reference's own kernels (tests/test_gpu_raymarching.py)."""
import torch
from torch.autograd import Function

from .. import _lib

_P, _S = _lib.ptr, _lib.stream_ptr


def _f32(t):
    return t.contiguous().float()


def march_rays_train(rays_o, rays_d, bound, density_grid, mean_density, iter_density, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False):
    """-> xyzs [m,3], dirs [m,3], deltas [m], rays [N,3] int32 (ray id, offset, count).  Sample slabs are handed out
    in order of arrival (atomic counter), as in the reference; M = N*1024 unless a running mean_count is given."""
    rays_o, rays_d = _f32(rays_o).view(-1, 3), _f32(rays_d).view(-1, 3)
    N, H, dev = rays_o.shape[0], density_grid.shape[0], rays_o.device
    M = N * 1024
    if not force_all_rays and mean_count > 0:
        M = mean_count + (align - mean_count % align if align > 0 else 0)
    xyzs, dirs = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev)
    deltas, rays = torch.zeros(M, device=dev), torch.empty(N, 3, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ac_march_rays_train(_P(rays_o), _P(rays_d), _P(_f32(density_grid)), float(mean_density), int(iter_density), float(bound),
                                              N, H, M, _P(xyzs), _P(dirs), _P(deltas), _P(rays), _P(step_counter), int(bool(perturb)), _S()),
               "march_rays_train")
    if force_all_rays or mean_count <= 0:
        m = int(step_counter[0].item())                       # the one D2H sync the reference has too (raymarching.py:55)
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


class _CompositeTrain(Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, bound):
        sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), rays.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        weights_sum, image = torch.empty(N, device=sigmas.device), torch.empty(N, 3, device=sigmas.device)
        _lib.check(_lib.lib().ac_composite_rays_train_forward(_P(sigmas), _P(rgbs), _P(deltas), _P(rays), float(bound), M, N, _P(weights_sum),
                                                              _P(image), _S()), "composite_rays_train_forward")
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.dims = (M, N, float(bound))
        return weights_sum, image

    @staticmethod
    def backward(ctx, g_ws, g_img):
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        M, N, bound = ctx.dims
        g_sig, g_rgb = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        _lib.check(_lib.lib().ac_composite_rays_train_backward(_P(_f32(g_ws)), _P(_f32(g_img)), _P(sigmas), _P(rgbs), _P(deltas), _P(rays),
                                                               _P(weights_sum), _P(image), bound, M, N, _P(g_sig), _P(g_rgb), _S()),
                   "composite_rays_train_backward")
        return g_sig, g_rgb, None, None, None


composite_rays_train = _CompositeTrain.apply


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_grid, mean_density, near, far, align=-1, perturb=False):
    rays_o, rays_d = _f32(rays_o).view(-1, 3), _f32(rays_d).view(-1, 3)
    H, dev = density_grid.shape[0], rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
    _lib.check(_lib.lib().ac_march_rays(n_alive, n_step, _P(rays_alive), _P(rays_t), _P(rays_o), _P(rays_d), float(bound), H,
                                        _P(_f32(density_grid)), float(mean_density), _P(_f32(near)), _P(_f32(far)), _P(xyzs), _P(dirs), _P(deltas),
                                        int(perturb), _S()), "march_rays")
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, normals, deltas, weights, depth, image, normal_map):
    """In place: accumulates into weights/depth/image/normal_map [N,...] and marks finished rays with rays_t = -1."""
    _lib.check(_lib.lib().ac_composite_rays(n_alive, n_step, _P(rays_alive), _P(rays_t), _P(_f32(sigmas)), _P(_f32(rgbs)), _P(_f32(normals)),
                                            _P(_f32(deltas)), _P(weights), _P(depth), _P(image), _P(normal_map), _S()), "composite_rays")


def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
    _lib.check(_lib.lib().ac_compact_rays(n_alive, _P(rays_alive), _P(rays_alive_old), _P(rays_t), _P(rays_t_old), _P(alive_counter), _S()),
               "compact_rays")
