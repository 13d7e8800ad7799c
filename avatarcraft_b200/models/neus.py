"""models.neus -- the legacy MLP NeuS model of the reference (models/neus.py: SDFNetwork :88, RenderingNetwork :243,
SingleVarianceNetwork :324, NeuSRenderer :333-770, build_neus :784), kept importable with the same constructor / call
signatures and result keys.

Scope: NOT on the accelerated path.  No reference entry point can run it (stylize.py:150-151 raises NotImplementedError
for --implicit_model neus, render_canonical.py / render_warp.py only construct instant_nsr), so it is plain torch here as
it is plain torch there -- 8 x 256 frequency-encoded MLPs are a library-GEMM workload, not the hash-grid hot path this
package exists for.  What it shares with models.instant_nsr is restated once: the NeuS section-alpha rule and the
deterministic importance sampler.
"""
from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..encoder import get_encoder


def _mlp_layers(owner, dims, weight_norm, init=None, skip_in=()):
    """lin0..linK on `owner` (the reference's attribute names = its state-dict keys)."""
    for l in range(len(dims) - 1):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        lin = nn.Linear(dims[l], out_dim)
        if init is not None:
            init(l, lin, out_dim)
        setattr(owner, f"lin{l}", nn.utils.weight_norm(lin) if weight_norm else lin)


class SDFNetwork(nn.Module):
    """Encoded xyz -> [sdf, geometry feature] (models/neus.py:88-241); geometric (sphere) initialisation, optional skip."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(), bias=0.5, scale=1, geometric_init=True, weight_norm=True,
                 inside_outside=False, use_tsdf: bool = False, encoder_type: str = None, encoder_config: dict = None,
                 use_fd: bool = False, use_id: bool = False):
        super().__init__()
        self.embed_fn_fine, input_ch = get_encoder(encoder_type or "frequency", encoder_config)
        dims = [input_ch + (3 if use_id else 0)] + [d_hidden] * n_layers + [d_out]
        self.num_layers, self.skip_in, self.scale = len(dims), tuple(skip_in), scale
        self.use_tsdf, self.use_fd, self.use_id = use_tsdf, use_fd, use_id
        last = self.num_layers - 2

        def init(l, lin, out_dim):
            if not geometric_init:
                return
            if l == last:
                sign = -1.0 if inside_outside else 1.0
                nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                nn.init.constant_(lin.bias, -sign * bias)
                return
            nn.init.constant_(lin.bias, 0.0)
            nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            if l == 0:
                nn.init.constant_(lin.weight[:, 3:], 0.0)
            elif l in self.skip_in:
                nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)

        _mlp_layers(self, dims, weight_norm, init, self.skip_in)
        self.activation = nn.Softplus(beta=100)

    def forward(self, inputs):
        inputs = inputs * self.scale
        enc = self.embed_fn_fine(inputs)
        inputs = torch.cat([inputs, enc], dim=-1) if self.use_id else enc
        x = inputs
        for l in range(self.num_layers - 1):
            if l in self.skip_in:
                x = torch.cat([x, inputs], 1) / np.sqrt(2)
            x = getattr(self, f"lin{l}")(x)
            if l < self.num_layers - 2:
                x = self.activation(x)
        sdf = x[:, :1] / self.scale
        return torch.cat([torch.tanh(sdf) if self.use_tsdf else sdf, x[:, 1:]], dim=-1)

    def sdf(self, x):
        return self.forward(x)[:, :1]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def finite_difference_normals_approximator(self, x, bound, epsilon=0.0005):
        """Central differences of the signed distance at +-epsilon per axis, points re-clamped to the bound (:205-222; the
        reference calls an undefined `forward_sdf` there -- `sdf` is what it means)."""
        cols = []
        for axis in range(3):
            e = torch.zeros(1, 3, device=x.device)
            e[0, axis] = epsilon
            cols.append(0.5 * (self.sdf((x + e).clamp(-bound, bound)) - self.sdf((x - e).clamp(-bound, bound))) / epsilon)
        return torch.cat(cols, dim=-1)

    def gradient(self, x):
        x.requires_grad_(True)
        with torch.enable_grad():
            y = self.sdf(x)
            g = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True, only_inputs=True)[0]
        return g.unsqueeze(1)


class RenderingNetwork(nn.Module):
    """IDR colour network (models/neus.py:243-321): modes 'idr' (points, view dirs, normals, features), 'no_view_dir',
    'no_normal'."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0, squeeze_out=True,
                 encoder_type: str = None, encoder_config: dict = None, activation: str = None):
        super().__init__()
        self.mode, self.squeeze_out, self.embedview_fn = mode, squeeze_out, None
        d0 = d_in + d_feature
        if mode == "no_view_dir":
            d0 -= 3
        else:
            self.embedview_fn, input_ch = get_encoder(encoder_type or "frequency", encoder_config)
            d0 += input_ch - 3
        dims = [d0] + [d_hidden] * n_layers + [d_out]
        self.num_layers = len(dims)
        _mlp_layers(self, dims, weight_norm)
        activation = activation or "relu"
        if activation not in ("relu", "softplus"):
            raise NotImplementedError(activation)
        self.activation = nn.ReLU() if activation == "relu" else nn.Softplus(beta=100)

    def forward(self, points, normals, view_dirs, feature_vectors):
        if self.embedview_fn is not None:
            view_dirs = self.embedview_fn(view_dirs)
        parts = {"idr": (points, view_dirs, normals, feature_vectors), "no_view_dir": (points, normals, feature_vectors),
                 "no_normal": (points, view_dirs, feature_vectors)}[self.mode]
        x = torch.cat(parts, dim=-1)
        for l in range(self.num_layers - 1):
            x = getattr(self, f"lin{l}")(x)
            if l < self.num_layers - 2:
                x = self.activation(x)
        return torch.sigmoid(x) if self.squeeze_out else x


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=x.device) * torch.exp(self.variance * 10.0)


def sample_pdf(bins, weights, n_samples, det=False, device=None):
    """Inverse-CDF sampling of the piecewise-constant pdf `weights` over `bins` (models/neus.py:52-85)."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    shape = list(cdf.shape[:-1]) + [n_samples]
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples, device=cdf.device).expand(shape)
    else:
        u = torch.rand(shape, device=cdf.device)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo, hi = (hi - 1).clamp(min=0), hi.clamp(max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    den = c_hi - c_lo
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return b_lo + (u - c_lo) / den * (b_hi - b_lo)


def _transmittance_weights(alpha):
    ones = torch.ones_like(alpha[:, :1])
    return alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-7], -1), -1)[:, :-1]


class NeuSRenderer(nn.Module):
    """Sphere-bounded NeuS renderer (models/neus.py:333-770): 64 coarse + 64 importance samples in 4 rounds, section alphas
    from the logistic CDF of the signed distance, optional NeRF++ outside model (`nerf`, n_outside > 0)."""

    def __init__(self, nerf, sdf_network, deviation_network, color_network, n_samples, n_importance, n_outside, up_sample_steps,
                 perturb):
        super().__init__()
        self.nerf, self.sdf_network, self.deviation_network, self.color_network = nerf, sdf_network, deviation_network, color_network
        self.n_samples, self.n_importance, self.n_outside = n_samples, n_importance, n_outside
        self.up_sample_steps, self.perturb = up_sample_steps, perturb

    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
        radius = torch.linalg.norm(pts, ord=2, dim=-1)
        inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
        s0, s1, z0, z1 = sdf[:, :-1], sdf[:, 1:], z_vals[:, :-1], z_vals[:, 1:]
        cos = (s1 - s0) / (z1 - z0 + 1e-5)
        prev = torch.cat([torch.zeros_like(cos[:, :1]), cos[:, :-1]], dim=-1)
        cos = torch.minimum(prev, cos).clip(-1e3, 0.0) * inside
        mid, half = (s0 + s1) * 0.5, cos * (z1 - z0) * 0.5
        c_prev, c_next = torch.sigmoid((mid - half) * inv_s), torch.sigmoid((mid + half) * inv_s)
        alpha = (c_prev - c_next + 1e-5) / (c_prev + 1e-5)
        return sample_pdf(z_vals, _transmittance_weights(alpha), n_importance, det=True).detach()

    def cat_z_vals(self, rays_o, rays_d, z_vals, new_z_vals, sdf, last=False):
        n, t = z_vals.shape
        z_vals, index = torch.sort(torch.cat([z_vals, new_z_vals], dim=-1), dim=-1)
        if not last:
            pts = rays_o[:, None, :] + rays_d[:, None, :] * new_z_vals[..., :, None]
            new_sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(n, -1)
            sdf = torch.gather(torch.cat([sdf, new_sdf], dim=-1), 1, index)
        return z_vals, sdf

    def render_core_outside(self, rays_o, rays_d, z_vals, sample_dist, nerf, background_rgb=None):
        """NeRF++ inverted-sphere background (models/neus.py:355-392)."""
        n, t = z_vals.shape
        dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], torch.full_like(z_vals[..., :1], sample_dist)], -1)
        mid = z_vals + dists * 0.5
        pts = rays_o[:, None, :] + rays_d[:, None, :] * mid[..., :, None]
        norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).clip(1.0, 1e10)
        pts = torch.cat([pts / norm, 1.0 / norm], dim=-1)
        dirs = rays_d[:, None, :].expand(n, t, 3)
        density, color = nerf(pts.reshape(-1, 4), dirs.reshape(-1, 3))
        alpha = (1.0 - torch.exp(-F.softplus(density.reshape(n, t)) * dists))
        weights = _transmittance_weights(alpha)
        color = torch.sigmoid(color).reshape(n, t, 3)
        out = (weights[:, :, None] * color).sum(dim=1)
        if background_rgb is not None:
            out = out + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        return {"color": out, "sampled_color": color, "alpha": alpha, "weights": weights}

    def render_core(self, rays_o, rays_d, z_vals, sample_dist, sdf_network, deviation_network, color_network, background_alpha=None,
                    background_sampled_color=None, background_rgb=None, cos_anneal_ratio=0.0):
        n, t = z_vals.shape
        dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1], torch.full_like(z_vals[..., :1], sample_dist)], -1)
        mid_z = z_vals + dists * 0.5
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
        dirs = rays_d[:, None, :].expand(n, t, 3).reshape(-1, 3)
        out = sdf_network(pts)
        sdf, feat = out[:, :1], out[:, 1:]
        grads = sdf_network.gradient(pts).squeeze()
        color = color_network(pts, grads, dirs, feat).reshape(n, t, 3)
        inv_s = deviation_network(torch.zeros([1, 3], device=pts.device))[:, :1].clip(1e-6, 1e6).expand(n * t, 1)
        cos = (dirs * grads).sum(-1, keepdim=True)
        iter_cos = -(F.relu(-cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-cos) * cos_anneal_ratio)
        half = iter_cos * dists.reshape(-1, 1) * 0.5
        c_prev, c_next = torch.sigmoid((sdf - half) * inv_s), torch.sigmoid((sdf + half) * inv_s)
        alpha = ((c_prev - c_next + 1e-5) / (c_prev + 1e-5)).reshape(n, t).clip(0.0, 1.0)
        radius = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(n, t)
        inside, relax = (radius < 1.0).float().detach(), (radius < 1.2).float().detach()
        if background_alpha is not None:
            alpha = torch.cat([alpha * inside + background_alpha[:, :t] * (1.0 - inside), background_alpha[:, t:]], dim=-1)
            color = torch.cat([color * inside[:, :, None] + background_sampled_color[:, :t] * (1.0 - inside)[:, :, None],
                               background_sampled_color[:, t:]], dim=1)
        weights = _transmittance_weights(alpha)
        image = (color * weights[:, :, None]).sum(dim=1)
        if background_rgb is not None:
            image = image + background_rgb * (1.0 - weights.sum(dim=-1, keepdim=True))
        err = (torch.linalg.norm(grads.reshape(n, t, 3), ord=2, dim=-1) - 1.0) ** 2
        return {"color": image, "sdf": sdf, "dists": dists, "gradients": grads.reshape(n, t, 3), "s_val": 1.0 / inv_s,
                "mid_z_vals": mid_z, "weights": weights, "cdf": c_prev.reshape(n, t),
                "gradient_error": (relax * err).sum() / (relax.sum() + 1e-5), "inside_sphere": inside}

    def render(self, rays_o: torch.Tensor, rays_d: torch.Tensor, near: float, far: float, perturb_overwrite=-1,
               n_importance_overwrite=-1, background_rgb=None, cos_anneal_ratio=0.0, render_can=False, posed_verts=None, faces=None,
               Ts=None):
        """rays [N,3] -> {'color_fine', 's_val', 'cdf_fine', 'weight_sum', 'weight_max', 'gradients', 'weights', 'gradient_error',
        'inside_sphere'} (models/neus.py:647-735).  As in the reference, importance sampling only runs when
        n_importance_overwrite > 0, and render_can / posed_verts / faces / Ts are accepted and unused."""
        n, dev = rays_o.shape[0], rays_o.device
        sample_dist = 2.0 / self.n_samples
        z_vals = near + (far - near) * torch.linspace(0.0, 1.0, self.n_samples, device=dev)[None, :]
        z_out = torch.linspace(1e-3, 1.0 - 1.0 / (self.n_outside + 1.0), self.n_outside, device=dev) if self.n_outside > 0 else None
        perturb = perturb_overwrite if perturb_overwrite >= 0 else self.perturb
        if perturb > 0:
            z_vals = z_vals + (torch.rand([n, 1], device=dev) - 0.5) * 2.0 / self.n_samples
            if z_out is not None:
                mids = 0.5 * (z_out[1:] + z_out[:-1])
                upper, lower = torch.cat([mids, z_out[-1:]], -1), torch.cat([z_out[:1], mids], -1)
                z_out = lower[None, :] + (upper - lower)[None, :] * torch.rand([n, z_out.shape[-1]], device=dev)
        if z_out is not None:
            z_out = far / torch.flip(z_out, dims=[-1]) + 1.0 / self.n_samples
        n_samples = self.n_samples
        if n_importance_overwrite > 0:
            with torch.no_grad():
                z_vals = z_vals.expand(n, self.n_samples).contiguous()
                pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[..., :, None]
                sdf = self.sdf_network.sdf(pts.reshape(-1, 3)).reshape(n, self.n_samples)
                for i in range(self.up_sample_steps):
                    new_z = self.up_sample(rays_o, rays_d, z_vals, sdf, self.n_importance // self.up_sample_steps, 64 * 2 ** i)
                    z_vals, sdf = self.cat_z_vals(rays_o, rays_d, z_vals, new_z, sdf, last=(i + 1 == self.up_sample_steps))
            n_samples = self.n_samples + self.n_importance
        z_vals = z_vals.expand(n, n_samples)
        bg_alpha = bg_color = None
        if self.n_outside > 0:
            z_feed, _ = torch.sort(torch.cat([z_vals, z_out.expand(n, -1)], dim=-1), dim=-1)
            outside = self.render_core_outside(rays_o, rays_d, z_feed, sample_dist, self.nerf)
            bg_color, bg_alpha = outside["sampled_color"], outside["alpha"]
        fine = self.render_core(rays_o, rays_d, z_vals, sample_dist, self.sdf_network, self.deviation_network, self.color_network,
                                background_rgb=background_rgb, background_alpha=bg_alpha, background_sampled_color=bg_color,
                                cos_anneal_ratio=cos_anneal_ratio)
        weights = fine["weights"]
        return {"color_fine": fine["color"], "s_val": fine["s_val"].reshape(n, n_samples).mean(dim=-1, keepdim=True),
                "cdf_fine": fine["cdf"], "weight_sum": weights.sum(dim=-1, keepdim=True),
                "weight_max": torch.max(weights, dim=-1, keepdim=True)[0], "gradients": fine["gradients"], "weights": weights,
                "gradient_error": fine["gradient_error"], "inside_sphere": fine["inside_sphere"]}

    @torch.no_grad()
    def extract_geometry(self, bound_min: List[float], bound_max: List[float], resolution: int, threshold: float = 0.0, device=None):
        """Iso-surface of -sdf on a `resolution`^3 lattice: the field is evaluated here, the surface extracted by this
        library's device kernel (ac_iso_surface), as models.instant_nsr.NeRFNetwork.extract_geometry does."""
        import ctypes
        from .. import _lib
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        lo, hi = np.asarray(bound_min, np.float32), np.asarray(bound_max, np.float32)
        axes = [torch.linspace(float(lo[i]), float(hi[i]), resolution, device=dev) for i in range(3)]
        u = torch.empty(resolution, resolution, resolution, device=dev)
        for i0 in range(0, resolution, 16):
            xs = axes[0][i0:i0 + 16]
            g = torch.stack(torch.meshgrid(xs, axes[1], axes[2], indexing="ij"), -1).reshape(-1, 3)
            u[i0:i0 + 16] = (-self.sdf_network.sdf(g)).reshape(len(xs), resolution, resolution)
        L, counter = _lib.lib(), torch.zeros(1, dtype=torch.int64, device=dev)

        def run(cap, pos, key):
            counter.zero_()
            _lib.check(L.ac_iso_surface(_lib.ptr(u), lo.ctypes.data_as(ctypes.c_void_p), hi.ctypes.data_as(ctypes.c_void_p), resolution,
                                        float(threshold), None if pos is None else _lib.ptr(pos), None if key is None else _lib.ptr(key),
                                        cap, _lib.ptr(counter), _lib.stream_ptr()), "ac_iso_surface")
            return int(counter.item())
        n_tri = run(0, None, None)
        if n_tri == 0:
            return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int64)
        pos = torch.empty(n_tri, 3, 3, device=dev); key = torch.empty(n_tri, 3, dtype=torch.int64, device=dev)
        run(n_tri, pos, key)
        uniq, inv = torch.unique(key.reshape(-1), return_inverse=True)
        verts = torch.empty(uniq.numel(), 3, device=dev)
        verts[inv] = pos.reshape(-1, 3)
        return verts.cpu().numpy(), inv.reshape(-1, 3).cpu().numpy()

    def freeze_module(self, module_name: str):
        for p in getattr(self, module_name).parameters():
            p.requires_grad = False


class OffsetNet(nn.Module):
    def __init__(self, pos_pe, neus):
        super().__init__()
        self.pos_pe, self.neus = pos_pe, neus

    def forward(self, input_pts, cur_iter=None):
        return self.neus(self.pos_pe(input_pts))


def build_neus(n_sdf: int = 6, n_color: int = 4, w_sdf: int = 256, w_color: int = 256, w_geo_feat: int = 256, xyz_encoder: str = None,
               dir_encoder: str = None, skip: list = None, rgb_activation: str = None, use_tsdf: bool = False, use_view_dir: bool = False,
               use_fd: bool = False, use_id: bool = False):
    """(NeuSRenderer, list of trainable parameters) with the reference's defaults (models/neus.py:784-886): frequency
    encoders (6 / 4 octaves), one skip at layer 4, variance 0.3, 64 + 64 samples in 4 importance rounds."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    pos_cfg = {"in_dim": 3, "freq_multires": 6, "hash_num_levels": 16, "hash_level_dim": 2, "hash_base_resolution": 16,
               "hash_per_level_scale": 1.3819, "hash_log2_hashmap_size": 19, "hash_desired_resolution": 2048}
    dir_cfg = {"in_dim": 3, "freq_multires": 4}
    sdf_net = SDFNetwork(d_out=w_geo_feat + 1, d_in=3, d_hidden=w_sdf, n_layers=n_sdf, skip_in=[4] if skip is None else skip, bias=0.5,
                         scale=1.0, geometric_init=True, weight_norm=True, use_tsdf=use_tsdf, encoder_type=xyz_encoder or "frequency",
                         encoder_config=pos_cfg, use_fd=use_fd, use_id=use_id).to(device)
    dev_net = SingleVarianceNetwork(init_val=0.3).to(device)
    color_net = RenderingNetwork(d_feature=w_geo_feat, mode="idr" if use_view_dir else "no_view_dir", d_in=9, d_out=3, d_hidden=w_color,
                                 n_layers=n_color, weight_norm=True, squeeze_out=True, encoder_type=dir_encoder or "frequency",
                                 encoder_config=dir_cfg, activation=rgb_activation).to(device)
    params = list(sdf_net.parameters()) + list(dev_net.parameters()) + list(color_net.parameters())
    neus = NeuSRenderer(None, sdf_net, dev_net, color_net, n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=0.0)
    return neus, params
