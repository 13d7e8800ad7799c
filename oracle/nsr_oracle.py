"""oracle/nsr_oracle.py -- TEST INFRASTRUCTURE ONLY.

Torch-CPU restatement of the reference's Instant-NSR render path: hash encode -> SDF MLP ->
hierarchical up-sampling -> finite-difference normals -> colour MLP -> NeuS alpha ->
front-to-back compositing.  It is the checker for the CUDA path and the "port" CPU
baseline of bench.py; it is never on the product path.

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is pinned by oracle/make_golden.py, which runs the reference's OWN
`models/instant_nsr.py` (imported from /root/reference, with oracle/hashgrid.py standing in
for the CUDA-only `_backend`) on the same seeded inputs and requires agreement; the
resulting vectors are committed under tests/golden/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import hashgrid as _hg


class _HashEncodeCPU(torch.autograd.Function):
    """Autograd wrapper of the C restatement (forward + table backward), mirroring
    encoder/hashencoder/hashgrid.py:11-73 so the oracle can produce parameter gradients."""

    @staticmethod
    def forward(ctx, x01, table, offsets, per_level_scale, base_resolution, scales):
        x01 = x01.detach().contiguous().float()
        B, D = x01.shape
        L, C = offsets.numel() - 1, table.shape[1]
        S = float(np.log2(per_level_scale))
        out = torch.empty(L, B, C)
        _hg.hash_encode_forward(x01, table.detach().contiguous(), offsets, out, B, D, C, L, S, base_resolution, False, None,
                                scales=scales)
        ctx.save_for_backward(x01, offsets)
        ctx.meta = (B, D, C, L, S, base_resolution, scales, table.shape)
        return out.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        x01, offsets = ctx.saved_tensors
        B, D, C, L, S, H, scales, shape = ctx.meta
        g = grad.reshape(B, L, C).permute(1, 0, 2).contiguous()
        gt = torch.zeros(shape)
        _hg.hash_encode_backward(g, x01, None, offsets, gt, B, D, C, L, S, H, False, None, None, scales=scales)
        return None, gt, None, None, None, None


def fold_weight_norm(g, v):
    """Effective weight of nn.utils.weight_norm(dim=0): g * v / ||v||_row
    (models/instant_nsr.py:555-556,585-586).  torch._weight_norm is the very function the
    reference's nn.utils.weight_norm hook evaluates, so the folded weights are bit-identical."""
    return torch._weight_norm(v, g, 0)


class OracleNSR:
    """Stateless functional model over a reference-layout state dict."""

    def __init__(self, state_dict, per_level_scale=None, base_resolution=16, training=False):
        sd = {k: v.detach().cpu().float() if v.is_floating_point() else v.detach().cpu()
              for k, v in state_dict.items()}
        self.table = sd["encoder.embeddings"].contiguous()
        self.offsets = sd["encoder.offsets"].to(torch.int32).contiguous()
        if per_level_scale is None:
            # models/instant_nsr.py:503-512 + hashgrid.py:84-86: desired_resolution 2048 overrides
            per_level_scale = float(np.exp2(np.log2(2048 / base_resolution) / (self.offsets.numel() - 2)))
        self.per_level_scale = per_level_scale
        self.base_resolution = base_resolution
        self.sdf_w = [fold_weight_norm(sd[f"sdf_net.{i}.weight_g"], sd[f"sdf_net.{i}.weight_v"]) for i in range(2)]
        self.sdf_b = [sd[f"sdf_net.{i}.bias"] for i in range(2)]
        self.col_w = [fold_weight_norm(sd[f"color_net.{i}.weight_g"], sd[f"color_net.{i}.weight_v"]) for i in range(3)]
        self.variance = sd["deviation_net.variance"]
        self.training = training
        self.level_scales = None  # optional override (e.g. read back from the GPU)

    # ---- encoder (encoder/hashencoder/hashgrid.py:126-142) -------------------------------
    def encode(self, x, bound, want_ids=False):
        x01 = (x + bound) / (2 * bound)
        if self.table.requires_grad and not want_ids:
            return _HashEncodeCPU.apply(x01, self.table, self.offsets, self.per_level_scale, self.base_resolution,
                                        self.level_scales)
        return _hg.encode(x01, self.table, self.offsets, self.per_level_scale, self.base_resolution,
                          want_ids=want_ids, scales=self.level_scales)

    # ---- SDF network (models/instant_nsr.py:627-642) --------------------------------------
    def forward_sdf(self, x, bound):
        h = torch.cat([x, self.encode(x, bound)], dim=-1)          # raw xyz is concatenated (include_input)
        h = F.softplus(F.linear(h, self.sdf_w[0], self.sdf_b[0]), beta=100)
        return F.linear(h, self.sdf_w[1], self.sdf_b[1])            # [:,0]=sdf, [:,1:]=15 geometry features

    # ---- colour network (models/instant_nsr.py:644-663, use_viewdirs=False branch) --------
    def forward_color(self, x, n, geo_feat):
        h = torch.cat([x, n, geo_feat], dim=-1)
        h = F.relu(F.linear(h, self.col_w[0]))
        h = F.relu(F.linear(h, self.col_w[1]))
        return torch.sigmoid(F.linear(h, self.col_w[2]))

    # ---- variance (models/instant_nsr.py:665-667,720-726) ---------------------------------
    def inv_s(self):
        return torch.exp(self.variance * 10.0).clip(1e-6, 1e6)

    # ---- central-difference gradient (models/instant_nsr.py:687-704) ----------------------
    def fd_gradient(self, x, bound, eps):
        cols = []
        for axis in range(3):
            step = torch.zeros(1, 3)
            step[0, axis] = eps
            hi = self.forward_sdf((x + step).clamp(-bound, bound), bound)[:, :1]
            lo = self.forward_sdf((x - step).clamp(-bound, bound), bound)[:, :1]
            cols.append(0.5 * (hi - lo) / eps)
        return torch.cat(cols, dim=-1)

    # ---- ray / box intersection (models/instant_nsr.py:58-77, 'cube') ---------------------
    @staticmethod
    def near_far(rays_o, rays_d, bound):
        t0 = (-bound - rays_o) / (rays_d + 1e-15)
        t1 = (bound - rays_o) / (rays_d + 1e-15)
        near = torch.where(t0 < t1, t0, t1).max(dim=-1, keepdim=True)[0].clamp(min=0.05)
        far = torch.where(t0 > t1, t0, t1).min(dim=-1, keepdim=True)[0]
        return near, far

    # ---- inverse-CDF sampling (models/instant_nsr.py:21-55, det=True) ---------------------
    @staticmethod
    def sample_pdf_det(bins, weights, n):
        w = weights + 1e-5
        pdf = w / w.sum(-1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
        u = torch.linspace(0.5 / n, 1.0 - 0.5 / n, steps=n).expand(cdf.shape[0], n).contiguous()
        hi = torch.searchsorted(cdf, u, right=True)
        lo = (hi - 1).clamp(min=0)
        hi = hi.clamp(max=cdf.shape[-1] - 1)
        c_lo, c_hi = torch.gather(cdf, 1, lo), torch.gather(cdf, 1, hi)
        b_lo, b_hi = torch.gather(bins, 1, lo), torch.gather(bins, 1, hi)
        den = c_hi - c_lo
        den = torch.where(den < 1e-5, torch.ones_like(den), den)
        return b_lo + (u - c_lo) / den * (b_hi - b_lo), (lo, hi)

    # ---- one importance round (models/instant_nsr.py:410-459) -----------------------------
    def up_sample(self, rays_o, rays_d, z, sdf, n_importance, inv_s, trace=None):
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., None]
        r = torch.linalg.norm(pts, ord=2, dim=-1)
        inside = (r[:, :-1] < 1.0) | (r[:, 1:] < 1.0)               # unit sphere, not `bound`
        z0, z1, s0, s1 = z[:, :-1], z[:, 1:], sdf[:, :-1], sdf[:, 1:]
        mid = (s0 + s1) * 0.5
        slope = (s1 - s0) / (z1 - z0 + 1e-5)
        prev = torch.cat([torch.zeros_like(slope[:, :1]), slope[:, :-1]], -1)
        slope = torch.minimum(prev, slope).clip(-1e3, 0.0) * inside
        d = z1 - z0
        c0 = torch.sigmoid((mid - slope * d * 0.5) * inv_s)
        c1 = torch.sigmoid((mid + slope * d * 0.5) * inv_s)
        alpha = (c0 - c1 + 1e-5) / (c0 + 1e-5)                        # NOT clipped here
        if trace is not None:
            trace["alpha"] = alpha
        trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        return self.sample_pdf_det(z, alpha * trans, n_importance)

    # ---- merge new depths (models/instant_nsr.py:461-475) ---------------------------------
    def cat_z_vals(self, rays_o, rays_d, z, z_new, sdf, bound, last):
        n, t = z.shape
        zz, order = torch.sort(torch.cat([z, z_new], -1), dim=-1)
        if not last:
            p = (rays_o[:, None, :] + rays_d[:, None, :] * z_new[..., None]).clamp(-bound, bound)
            s_new = self.forward_sdf(p.reshape(-1, 3), bound)[:, :1].reshape(n, -1)
            sdf = torch.gather(torch.cat([sdf, s_new], -1), 1, order)
        return zz, sdf, order

    # ---- the render core (models/instant_nsr.py:133-299), render_can=True branch ----------
    def enable_grad(self, state_dict):
        """Make the parameters autograd leaves (keys of the reference state dict) so `run_grad` yields
        d(loss)/d(parameter) exactly as the reference's autograd would (weight-norm fold included)."""
        self.params = {k: v.detach().clone().float().requires_grad_(True) for k, v in state_dict.items()
                       if v.is_floating_point()}
        p = self.params
        self.table = p["encoder.embeddings"]
        self.sdf_w = [fold_weight_norm(p[f"sdf_net.{i}.weight_g"], p[f"sdf_net.{i}.weight_v"]) for i in range(2)]
        self.sdf_b = [p[f"sdf_net.{i}.bias"] for i in range(2)]
        self.col_w = [fold_weight_norm(p[f"color_net.{i}.weight_g"], p[f"color_net.{i}.weight_v"]) for i in range(3)]
        self.variance = p["deviation_net.variance"]
        return self.params

    def run_grad(self, *args, **kwargs):
        """`run` with autograd enabled for the render core; sample placement stays gradient-free
        exactly as in the reference (`with torch.no_grad()`, models/instant_nsr.py:175-185)."""
        return self._run(*args, grad=True, **kwargs)

    @torch.no_grad()
    def run(self, *args, **kwargs):
        return self._run(*args, grad=False, **kwargs)

    def _run(self, rays_o, rays_d, num_steps, bound, upsample_steps, bg_color=None, cos_anneal_ratio=1.0,
             normal_epsilon_ratio=0.0, jitter=None, alpha_mask=None, trace=None, grad=False,
             verts=None, faces=None, Ts=None, use_mesh_guide=True):
        """rays_o/rays_d [N,3].  `jitter` [N,num_steps] in [0,1) replaces the reference's
        torch.rand draw (:162) so training-mode runs are reproducible.  Returns the same
        10-tuple as the reference.  `trace` (dict) receives intermediates for kernel tests."""
        rays_o = rays_o.reshape(-1, 3).float()
        rays_d = rays_d.reshape(-1, 3).float()
        N = rays_o.shape[0]
        near, far = self.near_far(rays_o, rays_d, bound)
        warp = None
        if verts is not None:                      # render_can=False branch (:147-153, :166-172, :198-203)
            from . import warp_oracle as _wo
            if use_mesh_guide:
                gn, gf = _wo.geometry_guided_near_far(rays_o, rays_d, verts, 0.05)
                near = torch.where(torch.isinf(gn)[:, None], near, gn[:, None])
                far = torch.where(torch.isinf(gf)[:, None], far, gf[:, None])

            def warp(p):                           # float64 numpy round trip, then .float() like the reference
                can, mask, *_ = _wo.warp_samples_to_canonical(p.numpy(), verts, faces, Ts, 0.05, device=getattr(self, 'warp_device', None))
                return torch.from_numpy(can), torch.from_numpy(mask)
        z = near + (far - near) * torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand(N, num_steps)
        sample_dist = (far - near) / num_steps
        if jitter is not None:
            z = z + (jitter - 0.5) * sample_dist
        pts = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
        if warp is not None:
            pts, _ = warp(pts)
        pts = pts.clamp(-bound, bound).float()
        T = num_steps
        if upsample_steps > 0:
            with torch.no_grad():           # sample placement carries no gradient (:175-185)
                sdf = self.forward_sdf(pts.reshape(-1, 3), bound)[:, :1].reshape(N, T)
                if trace is not None:
                    trace["coarse_z"], trace["coarse_sdf"] = z.clone(), sdf.clone()
                rounds = upsample_steps // 16
                for i in range(rounds):
                    z_new, bins = self.up_sample(rays_o, rays_d, z, sdf, 16, 64 * 2 ** i)
                    z, sdf, order = self.cat_z_vals(rays_o, rays_d, z, z_new, sdf, bound, last=(i + 1 == rounds))
                    if trace is not None:
                        trace[f"round{i}_znew"], trace[f"round{i}_bins"] = z_new.clone(), bins
                        trace[f"round{i}_z"], trace[f"round{i}_sdf"] = z.clone(), sdf.clone()
            T += upsample_steps
        # section mid-points (:187-206); NB the tail delta keeps the COARSE step count (:160)
        deltas = torch.cat([z[:, 1:] - z[:, :-1], sample_dist * torch.ones_like(z[:, :1])], -1)
        z_mid = torch.cat([z[:, :-1] + 0.5 * deltas[:, :-1], z[:, -1:]], -1)
        P = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_mid.unsqueeze(-1)
        if warp is not None:                       # NB cat_z_vals evaluated the up-sample SDF at UN-warped points (:464-469)
            P, wmask = warp(P)
            alpha_mask = wmask.float()
        P = P.clamp(-bound, bound).float().reshape(-1, 3)
        dirs = rays_d.unsqueeze(-2).expand(N, T, 3).reshape(-1, 3)
        out = self.forward_sdf(P, bound)
        sdf, feat = out[:, :1], out[:, 1:]
        grad = self.fd_gradient(P, bound, 0.005 * (1.0 - normal_epsilon_ratio))
        normal = grad / (1e-5 + torch.linalg.norm(grad, ord=2, dim=-1, keepdim=True))
        color = self.forward_color(P, normal, feat)
        inv_s = self.inv_s()
        cosv = (dirs * normal).sum(-1, keepdim=True)
        it = -(F.softplus(-cosv * 0.5 + 0.5, beta=100) * (1.0 - cos_anneal_ratio)
               + F.softplus(-cosv, beta=100) * cos_anneal_ratio)
        nxt = sdf + it * deltas.reshape(-1, 1) * 0.5
        prv = sdf - it * deltas.reshape(-1, 1) * 0.5
        c0, c1 = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        alpha = ((c0 - c1 + 1e-5) / (c0 + 1e-5)).reshape(N, T).clip(0.0, 1.0)
        if alpha_mask is not None:
            alpha = alpha * alpha_mask.reshape(N, T)
        weights = alpha * torch.cumprod(torch.cat([torch.ones(N, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
        wsum = weights.sum(-1, keepdim=True)
        color = color.reshape(N, T, 3)
        image = (color * weights[:, :, None]).sum(1)
        nmap = (normal.reshape(N, T, 3) * weights[:, :, None]).sum(1)
        depth = (weights * ((z - near) / (far - near)).clamp(0, 1)).sum(-1)
        pn = torch.linalg.norm(P, ord=2, dim=-1).reshape(N, T)
        relax = (pn < 1.2).float()
        gerr = (torch.linalg.norm(grad.reshape(N, T, 3), ord=2, dim=-1) - 1.0) ** 2
        eik = (relax * gerr).sum() / (relax.sum() + 1e-5)
        image = image + (1 - wsum) * (1 if bg_color is None else bg_color)
        if trace is not None:
            trace.update(z_mid=z_mid, sdf=sdf.reshape(N, T), grad=grad.reshape(N, T, 3), feat=feat.reshape(N, T, 15),
                         near=near, far=far, eik_num=(relax * gerr).sum(-1), eik_den=relax.sum(-1))
        return depth.reshape(1, N), weights, wsum, image.reshape(1, N, 3), nmap, eik, 0.0, color, alpha, z
