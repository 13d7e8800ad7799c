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
    net (training mode -> jitter) and the eval-mode frozen net_gt.  `jitter` ([n, 64] in [0,1), optional) replaces the
    per-patch random draw of the coarse samples (:161-162).

    With the flat-buffer optimiser and a white / black background the whole step runs on this library's kernels
    (`native_patch_step`: no autograd graph, no torch arithmetic); AC_TRAIN_STEP=autograd selects the autograd
    composition of the same ops (`autograd_patch_step`), which is also the path for any other optimiser."""
    import os
    if (isinstance(optimizer, FlatAdam) and bkg_key % 4 in (WHITE_BKG, 1) and os.environ.get("AC_TRAIN_STEP", "") != "autograd"):
        return native_patch_step(net_style, net_gt, optimizer, rays_o, rays_d, pixel_grad, batch_size, w_eikonal, use_opacity,
                                 bkg_key, rank, world, jitter)
    return autograd_patch_step(net_style, net_gt, optimizer, rays_o, rays_d, pixel_grad, batch_size, w_eikonal, use_opacity,
                               bkg_key, rank, world, jitter)


def autograd_patch_step(net_style, net_gt, optimizer, rays_o, rays_d, pixel_grad, batch_size=4096, w_eikonal=0.01,
                        use_opacity=True, bkg_key=WHITE_BKG, rank=0, world=1, jitter=None):
    """The step as a torch autograd composition of the fused ops (_SdfStencil, _ShadeComposite): the weight-norm fold, the
    loss arithmetic and the gradient accumulation are torch's."""
    n = rays_o.shape[0]
    batch_size = min(batch_size, n)
    optimizer.zero_grad()
    stats = {"eikonal": [], "opacity": []}
    for s, e, scale in shard_patches(n, batch_size, rank, world):
        o, d = rays_o[s:e], rays_d[s:e]
        if jitter is not None:
            out = net_style.render(o[None], d[None], num_steps=64, upsample_steps=64, bound=NSR_BOUND, staged=False,
                                   bg_color=None if bkg_key % 4 == WHITE_BKG else torch.zeros(e - s, 3, device=o.device),
                                   cos_anneal_ratio=1.0, normal_epsilon_ratio=0.0, render_can=True, perturb=True,
                                   jitter=jitter[s:e].contiguous())
            rgb, eik, extra = out["rgb"].reshape(-1, 3), out["gradient_error"], {"weight_sum": out["weight_sum"].reshape(-1, 1)}
        else:
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


class _NativeStepState:
    """Device buffers of native_patch_step that live as long as the model: gradient pointers into the optimiser's flat
    buffer, accumulators, device scalars."""

    def __init__(self, net, optimizer):
        named = dict(net.named_parameters())
        self.key = (optimizer.flat_grad.data_ptr(), tuple(p.data_ptr() for p in named.values()))
        g = lambda k: named[k].grad
        for p in named.values():
            if p.grad is None or not p.grad.is_contiguous():
                raise RuntimeError("native_patch_step needs the flat gradient views of FlatAdam (call optimizer.zero_grad() once)")
        self.named = named
        self.grad = {k: g(k) for k in named}
        dev = optimizer.flat_grad.device
        f32 = dict(device=dev, dtype=torch.float32)
        self.acc = torch.zeros(64 * 36 + 16 * 64, **f32)               # [dW0 | db0] (64 x 36) and dW1 (16 x 64), scaled
        self.scales = torch.zeros(3, **f32)
        self.g_eik = torch.zeros(1, **f32)
        self.opacity = torch.zeros(64, **f32)                          # one slot per patch of a step


def native_patch_step(net_style, net_gt, optimizer, rays_o, rays_d, pixel_grad, batch_size=4096, w_eikonal=0.01,
                      use_opacity=True, bkg_key=WHITE_BKG, rank=0, world=1, jitter=None, mark=None):
    """`mark(name)`: optional callback invoked on the stream after the patch loop / the all-reduce / the Adam launch.
    stylize.py:143-199 on this library's kernels only -- per patch (models/instant_nsr.py:133-299 with gradients):
         jitter (ac_fill_uniform) -> sample depths (ac_nsr_render, sampling only) -> section points -> 7-point SDF stencil
         forward -> shade forward (normals, colour MLP, alpha, compositing, eikonal) -> frozen net_gt opacity render ->
         shade backward (pixel gradient, eikonal and opacity terms seeded in-kernel) + ONE GEMM (colour weight gradients)
         -> SDF stencil backward (table scatter straight into the optimiser's flat gradient; weight gradients in TMEM)
         -> weight-norm backward of the five layers, added to the flat gradient
       then ONE all-reduce and ONE Adam launch.  No autograd graph, no torch arithmetic, no host synchronisation."""
    import ctypes
    from .. import _lib
    from ..models import instant_nsr as M
    if not isinstance(optimizer, FlatAdam):
        raise RuntimeError("native_patch_step needs utils.optim.FlatAdam")
    n = rays_o.shape[0]
    batch_size = min(batch_size, n)
    optimizer.zero_grad()
    st = getattr(net_style, "_native_step", None)
    named_now = tuple(p.data_ptr() for p in net_style.parameters())
    if st is None or st.key != (optimizer.flat_grad.data_ptr(), named_now):
        st = net_style._native_step = _NativeStepState(net_style, optimizer)
    L, sp = _lib.lib(), _lib.stream_ptr
    dev = rays_o.device
    num_steps, upsample_steps, bound, eps, car = 64, 64, float(NSR_BOUND), 0.005, 1.0
    stats = {"eikonal": [], "opacity": []}
    pixel_grad = pixel_grad.reshape(n, 3).float()
    G = st.grad
    patches = shard_patches(n, batch_size, rank, world)
    if use_opacity and net_gt is not None:
        if len(patches) > st.opacity.numel():
            st.opacity = torch.zeros(len(patches), device=dev, dtype=torch.float32)
        _lib.check(L.ac_zero(_lib.ptr(st.opacity), st.opacity.numel() * 4, sp()), "ac_zero")
    # Everything that does not depend on the pixel gradient or on the patch order is done ONCE for all of this rank's rays: the
    # parameters only change at optimizer.step(), so the sampled depths of every patch (and the frozen net's opacities) are
    # the same whether computed patch by patch (the reference, stylize.py:153-181) or in one launch -- and a 4096-ray
    # sampling launch is latency bound (one ray per warp), 16 of them cost 8x one 65 536-ray launch.
    if not patches:                       # more ranks than rays: this rank only joins the collectives
        optimizer.all_reduce()
        optimizer.step()
        return _LazyStats(stats)
    if len(patches) == 1:
        idx = None
        o_all, d_all = rays_o[patches[0][0]:patches[0][1]].float().contiguous(), rays_d[patches[0][0]:patches[0][1]].float().contiguous()
    else:
        idx = torch.cat([torch.arange(s, e, device=dev) for s, e, _ in patches]) if world > 1 else None
        o_all = (rays_o if idx is None else rays_o[idx]).float().contiguous()
        d_all = (rays_d if idx is None else rays_d[idx]).float().contiguous()
    n_all = o_all.shape[0]
    if jitter is not None:
        first = patches[0][0] if len(patches) == 1 else 0
        jit_all = (jitter[first:first + n_all] if idx is None else jitter[idx]).float().contiguous()
    else:
        jit_all = torch.empty(n_all, num_steps, device=dev, dtype=torch.float32)
        # keyed by a draw from torch's CPU generator: reproducible under torch.manual_seed, restored by a resume
        # (utils/checkpoint.py saves the generator state), and no device work or synchronisation
        seed = (int(torch.randint(0, 2 ** 62, (1,)).item()) + rank * 0x9E3779B1) & 0x7FFFFFFFFFFFFFFF
        _lib.check(L.ac_fill_uniform(_lib.ptr(jit_all), n_all * num_steps, seed, sp()), "ac_fill_uniform")
    z_all = net_style._sample_depths(o_all, d_all, num_steps, upsample_steps, bound, jit_all)
    wsum_gt_all = None
    if use_opacity and net_gt is not None:
        with torch.no_grad():
            wsum_gt_all = net_gt.run(o_all[None], d_all[None], num_steps, bound, upsample_steps, None, 1.0, 0.0, per_sample_outputs=False,
                                     opacity_only=True)[2].reshape(-1)
    at = 0
    for ip, (s, e, scale) in enumerate(patches):
        m = e - s
        o, d, z = o_all[at:at + m], d_all[at:at + m], z_all[at:at + m]
        wsum_gt = None if wsum_gt_all is None else wsum_gt_all[at:at + m]
        at += m
        P = net_style._section_points(o, d, z, bound)
        Mp = P.shape[0]
        centre = torch.empty(Mp, 16, device=dev, dtype=torch.float32)
        fd = torch.empty(6, Mp, device=dev, dtype=torch.float32)
        model = net_style._device_model()
        fc = M._feature_cache(net_style, 7 * Mp, dev)          # encoded features of the 7 Mp points, reused by the backward below
        _lib.check(L.ac_nsr_forward_sdf_stencil_cache(ctypes.byref(model), _lib.ptr(P), Mp, bound, eps, _lib.ptr(centre), _lib.ptr(fd),
                                                      _lib.ptr(fc), fc.numel(), sp()), "ac_nsr_forward_sdf_stencil_cache")
        bg = None if bkg_key % 4 == WHITE_BKG else torch.zeros(m, 3, device=dev)
        bufs = M._shade_forward(net_style, o, d, z, P, centre, fd, bg, num_steps, bound, eps, car)
        g_eik = None
        if w_eikonal > 0.0:
            if scale != 1.0:        # a split patch: this rank's share of the masked mean (one scalar all-reduce)
                share = masked_mean_share(bufs["eik_out"][1], group=optimizer.group)
                st.g_eik.copy_((share * w_eikonal).reshape(1))
                stats["eikonal"].append((bufs["eik_out"][0:1], share))
            else:
                if getattr(st, "g_eik_value", None) != w_eikonal:
                    st.g_eik.fill_(w_eikonal); st.g_eik_value = w_eikonal
                stats["eikonal"].append((bufs["eik_out"][0:1], 1.0))
            if scale != 1.0:
                st.g_eik_value = None
            g_eik = st.g_eik
        g_centre, g_fd, C, cscale = M._shade_backward(net_style, o, d, z, P, centre, fd, bg, num_steps, bound, eps, car, bufs,
                                                      pixel_grad[s:e], g_eik=g_eik, wsum_gt=wsum_gt, opacity_weight=1e5 * scale,
                                                      g_variance=G["deviation_net.variance"], g_b1=G["sdf_net.1.bias"],
                                                      opacity_loss=st.opacity[ip:ip + 1] if wsum_gt is not None else None)
        if wsum_gt is not None:
            stats["opacity"].append((st.opacity[ip:ip + 1], 1.0 / scale))
        _lib.check(L.ac_nsr_sdf_backward_scales(ctypes.byref(model), _lib.ptr(g_centre), _lib.ptr(g_fd), Mp, _lib.ptr(st.scales), sp()),
                   "ac_nsr_sdf_backward_scales")
        _lib.check(L.ac_zero(_lib.ptr(st.acc), st.acc.numel() * 4, sp()), "ac_zero")
        acc0, acc1 = st.acc[:64 * 36], st.acc[64 * 36:]
        ws = M._backward_workspace(net_style, 7 * Mp, dev)
        _lib.check(L.ac_nsr_sdf_backward_stencil_ws(ctypes.byref(model), _lib.ptr(P), Mp, bound, eps, _lib.ptr(g_centre), _lib.ptr(g_fd),
                                                    _lib.ptr(st.scales), _lib.ptr(G["encoder.embeddings"]), _lib.ptr(acc0), _lib.ptr(acc1),
                                                    _lib.ptr(ws), ws.numel(), _lib.ptr(fc), sp()), "ac_nsr_sdf_backward_stencil_ws")
        sdf, col, P_ = net_style.sdf_net, net_style.color_net, _lib.ptr
        s0, s1 = st.scales[0:1], st.scales[1:2]

        def layer(dW, ld, lin, prefix, scale_t, db=None, db_col=0):
            return _lib.WeightNormLayer(dW=dW, v=lin.weight_v.data_ptr(), g=lin.weight_g.data_ptr(), dv=G[prefix + ".weight_v"].data_ptr(),
                                        dg=G[prefix + ".weight_g"].data_ptr(), rows=lin.weight_v.shape[0], cols=lin.weight_v.shape[1], ldw=ld,
                                        scale=scale_t.data_ptr(), db=None if db is None else db.data_ptr(), db_col=db_col)
        layers = (_lib.WeightNormLayer * 5)(
            layer(acc0.data_ptr(), 36, sdf[0], "sdf_net.0", s0, G["sdf_net.0.bias"], 35),
            layer(acc1.data_ptr(), 64, sdf[1], "sdf_net.1", s1),
            layer(C.data_ptr() + 4 * (64 * 160 + 64), 160, col[0], "color_net.0", cscale),
            layer(C.data_ptr(), 160, col[1], "color_net.1", cscale),
            layer(C.data_ptr() + 4 * (128 * 160 + 96), 160, col[2], "color_net.2", cscale))
        _lib.check(L.ac_nsr_weight_norm_backward(layers, 5, sp()), "ac_nsr_weight_norm_backward")
    if mark is not None:          # bench.py: CUDA events between the phases of the step
        mark("pass2")
    optimizer.all_reduce()
    if mark is not None:
        mark("allreduce")
    optimizer.step()
    if mark is not None:
        mark("adam")
    return _LazyStats(stats)


class _LazyStats(dict):
    """Per-patch device scalars of a native step; the means are formed (a few tiny torch ops) only when a key is read, so a
    step whose statistics nobody looks at launches nothing for them."""

    def __init__(self, parts):
        super().__init__()
        self._parts = parts

    def __getitem__(self, k):
        v = self._parts[k]
        return torch.cat([t.reshape(1) * w for t, w in v]).mean() if v else None

    def get(self, k, default=None):
        return self[k] if k in self._parts else default

    def keys(self):
        return self._parts.keys()

    def items(self):
        return [(k, self[k]) for k in self._parts]

    def __iter__(self):
        return iter(self._parts)

    def __len__(self):
        return len(self._parts)

    def __contains__(self, k):
        return k in self._parts
