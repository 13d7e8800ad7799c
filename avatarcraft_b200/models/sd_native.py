"""Native (sm_100a) forward of the Stable-Diffusion UNet for the SDS step: walks a `UNet2DConditionModel`'s parameters
and evaluates it with the kernels of csrc/sd_ops.cu -- TMA-fed tcgen05 GEMMs, implicit-GEMM 3x3 convolutions, a fused
attention kernel, and GroupNorm/SiLU, LayerNorm, GEGLU producers of the fp16 operands.  Same arithmetic graph as `UNet2DConditionModel.forward` (the autograd / torch-op path),
with fp16 GEMM operands and fp32 accumulation; tests/test_gpu_sds.py compares the two on identical weights.

Layout: activations are contiguous fp32 [B, H, W, C] (== the [B*H*W, C] token matrix); fp16 copies of the weights
(conv kernels permuted to [N, ky, kx, C]) are cached per parameter and refreshed when the parameter's version changes.
No fallback: every op raises if libavatarcraft_b200.so is missing or a tensor is not on a CUDA device."""
import ctypes
import os

import torch
import torch.nn.functional as F

from torch.utils.weak import WeakIdKeyDictionary

from .. import _lib
from .sd_blocks import timestep_embedding

_W16 = WeakIdKeyDictionary()
IMPLICIT_CONV = True   # 3x3 stride-1 convs as implicit GEMMs (TMA-shifted activation windows) instead of im2col + GEMM
FLASH = True          # fused attention kernel; False = Q K^T GEMM -> softmax -> P V GEMM (kept for A/B and head dims > 128)


def _check(rc, what):
    _lib.check(rc, what)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _w16(param, kind):
    """fp16 GEMM operand of a weight: 'linear' [N,K]; 'conv1' [N,C,1,1] -> [N,C]; 'conv3' [N,C,3,3] -> [N,3,3,Cp].
    Cached per parameter OBJECT (weak reference: a new module whose parameter happens to reuse a dead one's id() and
    device address can never hit a stale entry) and refreshed when the parameter's storage or version changes."""
    slot = _W16.get(param)
    if slot is None:
        slot = {}
        _W16[param] = slot
    hit = slot.get(kind)
    if hit is not None and hit[0] == (param.data_ptr(), param._version, tuple(param.shape)):
        return hit[1]
    w = param.detach()
    if kind == "conv3":
        N, C = w.shape[0], w.shape[1]
        Cp = (C + 7) // 8 * 8
        buf = torch.zeros(N, 3, 3, Cp, device=w.device, dtype=torch.float16)
        buf[..., :C] = w.permute(0, 2, 3, 1)
        w16 = buf.reshape(N, 9 * Cp)
    elif kind == "conv1":
        w16 = w.reshape(w.shape[0], w.shape[1]).to(torch.float16).contiguous()
    else:
        w16 = w.to(torch.float16).contiguous()
    if w16.shape[1] % 8:                       # operand rows must be 16 B multiples
        w16 = F.pad(w16, (0, 8 - w16.shape[1] % 8)).contiguous()
    slot[kind] = ((param.data_ptr(), param._version, tuple(param.shape)), w16)
    return w16


def gemm(A16, W16, M, N, K, bias=None, group_bias=None, rows_per_group=0, residual=None, out_f16=False, out=None,
         lda=None, ldw=None, ldc=None, batch=(1, 1), sA=(0, 0), sW=(0, 0), sC=(0, 0)):
    """Thin wrapper of ac_sd_gemm_f16 (see include/avatarcraft_b200.h)."""
    dev = A16.device
    if out is None:
        out = torch.empty(M, N, device=dev, dtype=torch.float16 if out_f16 else torch.float32)
    lda = A16.stride(-2) if lda is None else lda
    ldw = W16.stride(-2) if ldw is None else ldw
    ldc = N if ldc is None else ldc
    _check(_lib.lib().ac_sd_gemm_f16(_p(A16), _p(W16), _p(bias), _p(group_bias), int(rows_per_group), _p(residual), _p(out), int(out_f16),
                                     int(M), int(N), int(K), int(lda), int(ldw), int(ldc), int(N if residual is not None else 0),
                                     int(batch[0]), int(batch[1]), int(sA[0]), int(sA[1]), int(sW[0]), int(sW[1]), int(sC[0]), int(sC[1]),
                                     _lib.stream_ptr()), "ac_sd_gemm_f16")
    return out


def cast16(x):
    x = x.contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _check(_lib.lib().ac_sd_cast_f16(_p(x), x.numel(), _p(out), _lib.stream_ptr()), "ac_sd_cast_f16")
    return out


def gn_stats(x, groups, eps):
    """x [B,H,W,C] fp32 -> stats [B,G,2] (mean, rstd)."""
    B, H, W, C = x.shape
    ws = torch.empty(2 * B * groups, device=x.device, dtype=torch.float64)
    stats = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
    _check(_lib.lib().ac_sd_group_norm_stats(_p(x), B, H * W, C, groups, float(eps), _p(ws), _p(stats), _lib.stream_ptr()), "ac_sd_group_norm_stats")
    return stats


def im2col(x, ksize, stride=1, pad=0, up=False, Ho=None, Wo=None, norm=None):
    """x [B,Hs,Ws,C] fp32 -> fp16 [B*Ho*Wo, k*k*Cp]; norm = (GroupNormAct module) applies GroupNorm(+SiLU) on the fly."""
    B, Hs, Ws, C = x.shape
    Hin, Win = (2 * Hs, 2 * Ws) if up else (Hs, Ws)
    Ho = Hin if Ho is None else Ho
    Wo = Win if Wo is None else Wo
    Cp = (C + 7) // 8 * 8
    out = torch.empty(B * Ho * Wo, ksize * ksize * Cp, device=x.device, dtype=torch.float16)
    st = gam = bet = None
    G = act = 0
    if norm is not None:
        st = gn_stats(x, norm.num_groups, norm.eps)
        gam, bet, G, act = norm.weight.detach(), norm.bias.detach(), norm.num_groups, int(norm.act)
    _check(_lib.lib().ac_sd_im2col_f16(_p(x), B, Hs, Ws, C, ksize, stride, pad, int(up), Ho, Wo, _p(st), _p(gam), _p(bet), G, act, _p(out),
                                       _lib.stream_ptr()), "ac_sd_im2col_f16")
    return out, Ho, Wo


def _tile_ok(H, W):
    """A 128-pixel GEMM tile must be whole image rows (see ac_sd_conv3x3_f16).  Restricted to square maps: the UNet's only
    case and the one validated on the GPU."""
    if H != W or W > 128 or 128 % W:
        return False
    bh = 128 // W
    return (H % bh == 0) if bh <= H else (bh % H == 0)


def conv(x, mod, norm=None, up=False, group_bias=None, residual=None):
    """Conv2d module `mod` (1x1 or 3x3, stride 1/2, padding 0/1) on NHWC x, optionally fused with a preceding
    GroupNorm(+SiLU), a nearest x2 up-sampling, a per-(batch, channel) bias and a residual add.  -> [B,Ho,Wo,N] fp32."""
    B, Hs, Ws, C = x.shape
    k, stride, pad = mod.kernel_size[0], mod.stride[0], mod.padding[0]
    Hin, Win = (2 * Hs, 2 * Ws) if up else (Hs, Ws)
    Ho = (Hin + 2 * pad - k) // stride + 1
    Wo = (Win + 2 * pad - k) // stride + 1
    if IMPLICIT_CONV and k == 3 and stride == 1 and pad == 1 and not up and C % 64 == 0 and _tile_ok(Hs, Ws):
        # implicit GEMM: the (optionally normalised) activation is written once as fp16 NHWC and the nine shifted windows
        # are fetched by TMA inside the GEMM -- no [M, 9C] im2col buffer
        act16, _, _ = im2col(x, 1, norm=norm)                                           # [B*Hs*Ws, C] fp16 == NHWC
        N = mod.out_channels
        out = torch.empty(B, Hs, Ws, N, device=x.device, dtype=torch.float32)
        _check(_lib.lib().ac_sd_conv3x3_f16(_p(act16), _p(_w16(mod.weight, "conv3")), _p(None if mod.bias is None else mod.bias.detach()),
                                            _p(group_bias), _p(None if residual is None else residual.reshape(-1, N)), _p(out), B, Hs, Ws, C, N,
                                            _lib.stream_ptr()), "ac_sd_conv3x3_f16")
        return out
    A, Ho, Wo = im2col(x, k, stride, pad, up, Ho, Wo, norm)
    W16 = _w16(mod.weight, "conv3" if k == 3 else "conv1")
    N, K = mod.out_channels, A.shape[1]
    res = None if residual is None else residual.reshape(-1, N)
    out = gemm(A, W16, B * Ho * Wo, N, K, bias=None if mod.bias is None else mod.bias.detach(), group_bias=group_bias,
               rows_per_group=Ho * Wo, residual=res)
    return out.reshape(B, Ho, Wo, N)


def linear(x16, mod, M, residual=None, out_f16=False):
    """Linear module on an fp16 operand [M, K] -> [M, N]."""
    W16 = _w16(mod.weight, "linear")
    return gemm(x16, W16, M, mod.out_features, mod.in_features, bias=None if mod.bias is None else mod.bias.detach(), residual=residual,
                out_f16=out_f16)


def layer_norm16(x, mod):
    M, C = x.shape
    out = torch.empty(M, C, device=x.device, dtype=torch.float16)
    _check(_lib.lib().ac_sd_layer_norm_f16(_p(x), M, C, _p(mod.weight.detach()), _p(mod.bias.detach()), float(mod.eps), _p(out), _lib.stream_ptr()),
           "ac_sd_layer_norm_f16")
    return out


def attention(xq16, ctx16, attn, B, Lq, Lk, residual):
    """Multi-head attention of module `attn` (to_q/to_k/to_v/to_out.0): xq16 [B*Lq, Cq] fp16, ctx16 [B*Lk, Cc] fp16
    -> to_out(softmax(q k^T * scale) v) + residual, fp32 [B*Lq, Cq].  V is produced already transposed ([B, inner, Lk]) by
    swapping the value projection's GEMM operands; head dims <= 128 go through the fused attention kernel (scores stay
    on the SM), larger ones through two batched GEMMs around the softmax kernel."""
    dev = xq16.device
    heads = attn.heads
    inner = attn.to_q.out_features
    d = inner // heads
    Cc = ctx16.shape[1]
    Lp = (Lk + 7) // 8 * 8
    q = linear(xq16, attn.to_q, B * Lq, out_f16=True)                                   # [B*Lq, inner]
    k = linear(ctx16, attn.to_k, B * Lk, out_f16=True)                                  # [B*Lk, inner]
    vT = torch.zeros(B, inner, Lp, device=dev, dtype=torch.float16) if Lp != Lk else torch.empty(B, inner, Lp, device=dev, dtype=torch.float16)
    gemm(_w16(attn.to_v.weight, "linear"), ctx16, inner, Lk, Cc, out_f16=True, out=vT, ldc=Lp, batch=(B, 1), sA=(0, 0), sW=(Lk * Cc, 0),
         sC=(inner * Lp, 0))
    if FLASH and d % 8 == 0 and d <= 128:
        # fused: scores and probabilities never leave the SM (csrc/sd_ops.cu: sd_flash_attn_kernel)
        o16 = torch.empty(B * Lq, inner, device=dev, dtype=torch.float16)
        _check(_lib.lib().ac_sd_flash_attention_f16(_p(q), _p(k), _p(vT), _p(o16), B, heads, Lq, Lk, d, inner, inner, Lp, inner, float(attn.scale),
                                                    _lib.stream_ptr()), "ac_sd_flash_attention_f16")
        return linear(o16, attn.to_out[0], B * Lq, residual=residual)
    scores = torch.empty(B, heads, Lq, Lp, device=dev, dtype=torch.float32)
    gemm(q, k, Lq, Lk, d, out=scores, lda=inner, ldw=inner, ldc=Lp, batch=(B, heads), sA=(Lq * inner, d), sW=(Lk * inner, d),
         sC=(heads * Lq * Lp, Lq * Lp))
    probs = torch.empty(B, heads, Lq, Lp, device=dev, dtype=torch.float16)
    _check(_lib.lib().ac_sd_softmax_f16(_p(scores), B * heads * Lq, Lk, Lp, Lp, float(attn.scale), _p(probs), _lib.stream_ptr()), "ac_sd_softmax_f16")
    o16 = torch.empty(B * Lq, inner, device=dev, dtype=torch.float16)
    gemm(probs, vT, Lq, d, Lk, out_f16=True, out=o16, lda=Lp, ldw=Lp, ldc=inner, batch=(B, heads), sA=(heads * Lq * Lp, Lq * Lp),
         sW=(inner * Lp, d * Lp), sC=(Lq * inner, d))
    return linear(o16, attn.to_out[0], B * Lq, residual=residual)


def transformer(x, mod, ctx16, B_ctx_len):
    """Transformer2DModel on NHWC x [B,H,W,C] -> same shape."""
    B, H, W, C = x.shape
    L, M = H * W, B * H * W
    A, _, _ = im2col(x, 1, norm=mod.norm)                                            # GroupNorm (no act) -> fp16 [M, C]
    pin = mod.proj_in
    Wp = _w16(pin.weight, "linear" if mod.use_linear else "conv1")
    h = gemm(A, Wp, M, Wp.shape[0], C, bias=pin.bias.detach())                           # [M, inner] fp32
    for blk in mod.transformer_blocks:
        n1 = layer_norm16(h, blk.norm1)
        h = attention(n1, n1, blk.attn1, B, L, L, residual=h)
        n2 = layer_norm16(h, blk.norm2)
        h = attention(n2, ctx16, blk.attn2, B, L, B_ctx_len, residual=h)
        n3 = layer_norm16(h, blk.norm3)
        u = linear(n3, blk.ff.net[0].proj, M)                                            # [M, 8C] fp32
        inner_ff = u.shape[1] // 2
        g16 = torch.empty(M, inner_ff, device=x.device, dtype=torch.float16)
        _check(_lib.lib().ac_sd_geglu_f16(_p(u), M, inner_ff, _p(g16), _lib.stream_ptr()), "ac_sd_geglu_f16")
        h = linear(g16, blk.ff.net[2], M, residual=h)
    pout = mod.proj_out
    Wo = _w16(pout.weight, "linear" if mod.use_linear else "conv1")
    out = gemm(cast16(h), Wo, M, C, h.shape[1], bias=pout.bias.detach(), residual=x.reshape(M, C))
    return out.reshape(B, H, W, C)


def resnet(x, mod, temb_act16):
    """ResnetBlock2D on NHWC: conv1(silu(norm1 x)) + time projection -> conv2(silu(norm2 .)) + shortcut(x)."""
    B = x.shape[0]
    gb = None
    if mod.time_emb_proj is not None and temb_act16 is not None:
        gb = linear(temb_act16, mod.time_emb_proj, B)                                    # [B, Cout] fp32, added in conv1's epilogue
    h = conv(x, mod.conv1, norm=mod.norm1, group_bias=gb)
    sc = x if mod.conv_shortcut is None else conv(x, mod.conv_shortcut)
    return conv(h, mod.conv2, norm=mod.norm2, residual=sc)


GRAPH = os.environ.get("AC_SD_GRAPH", "1") != "0"     # replay the UNet forward as one CUDA graph (same kernels, no launch gaps)
GRAPH_WARMUP = 2                                      # eager calls per input shape before the capture


class _UNetGraph:
    """One captured forward for one input signature: static input / output buffers + the graph."""

    def __init__(self, unet, sample, timestep, ctx):
        self.sample = sample.detach().clone()
        self.t = torch.as_tensor(timestep, device=sample.device).reshape(-1).clone()
        self.ctx = ctx.detach().clone()
        self.graph = torch.cuda.CUDAGraph()
        count = _lib.lib().ac_launch_count
        torch.cuda.synchronize(sample.device)
        n0 = count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):    # other threads (NCCL watchdog) may touch CUDA meanwhile
            self.out = _unet_forward_eager(unet, self.sample, self.t, self.ctx)
        self.launches = int(count() - n0)             # kernels of this library inside the graph (ac_launch_count bookkeeping)

    def __call__(self, sample, timestep, ctx):
        self.sample.copy_(sample)
        self.t.copy_(torch.as_tensor(timestep, device=self.sample.device).reshape(-1))
        self.ctx.copy_(ctx)
        self.graph.replay()
        _lib.lib().ac_launch_count_add(self.launches)
        return self.out.clone()                       # [B,4,64,64]: the caller may keep it across calls


def _weights_stamp(unet):
    """Changes whenever a parameter is re-assigned or written in place (the graph holds fp16 copies of the weights)."""
    plist = unet.__dict__.get("_native_param_list")          # walking the module tree costs 1.9 ms per call; the list is kept
    if plist is None:                                        # (in-place loads / .to() keep the Parameter objects)
        plist = unet.__dict__["_native_param_list"] = list(unet.parameters())
    s = 0
    for p in plist:
        s += p._version + (p.data_ptr() & 0xFFFFFF)
    return s


def unet_forward(unet, sample, timestep, encoder_hidden_states):
    """UNet2DConditionModel.forward on the native kernels: sample [B,Cin,H,W] -> [B,Cout,H,W] fp32.
    After GRAPH_WARMUP eager calls with one input signature the launch sequence (~700 kernels of this library and a few torch
    copies) is captured once and replayed: the kernels are identical, the CPU launch cost and the gaps between the small
    kernels go away.  The result is copied out of the graph's output buffer, so it stays valid across calls.
    AC_SD_GRAPH=0 keeps every call eager."""
    if not sample.is_cuda:
        raise RuntimeError("sd_native.unet_forward needs CUDA tensors (no CPU path)")
    if not GRAPH or torch.cuda.is_current_stream_capturing():
        return _unet_forward_eager(unet, sample, timestep, encoder_hidden_states)
    t = torch.as_tensor(timestep, device=sample.device)
    key = (tuple(sample.shape), sample.dtype, int(t.numel()), t.dtype, tuple(encoder_hidden_states.shape), encoder_hidden_states.dtype,
           sample.device.index)
    cache = unet.__dict__.setdefault("_native_graphs", {})
    stamp = _weights_stamp(unet)
    if cache.get("stamp") != stamp:
        cache.clear(); cache["stamp"] = stamp
    slot = cache.get(key)
    if isinstance(slot, _UNetGraph):
        return slot(sample, timestep, encoder_hidden_states)
    seen = 0 if slot is None else slot
    if seen >= GRAPH_WARMUP:
        try:
            g = cache[key] = _UNetGraph(unet, sample, timestep, encoder_hidden_states)
            return g(sample, timestep, encoder_hidden_states)
        except Exception as e:                        # capture refused (e.g. an allocation inside it): stay eager, say so once
            import warnings
            warnings.warn(f"sd_native: CUDA graph capture of the UNet forward failed ({type(e).__name__}: {e}); running eagerly")
            cache[key] = -(1 << 30)
            return _unet_forward_eager(unet, sample, timestep, encoder_hidden_states)
    cache[key] = seen + 1
    return _unet_forward_eager(unet, sample, timestep, encoder_hidden_states)


@torch.no_grad()
def _unet_forward_eager(unet, sample, timestep, encoder_hidden_states):
    cfg = unet.config
    B = sample.shape[0]
    with torch.cuda.device(sample.device):
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(B)
        emb16 = cast16(timestep_embedding(t, cfg.block_out_channels[0]))
        te = unet.time_embedding
        temb = linear(cast16(F.silu(linear(emb16, te.linear_1, B))), te.linear_2, B)
        temb_act16 = cast16(F.silu(temb))
        ctx = encoder_hidden_states.float().contiguous()
        Lc = ctx.shape[1]
        ctx16 = cast16(ctx.reshape(B * Lc, -1))
        h = conv(sample.float().permute(0, 2, 3, 1).contiguous(), unet.conv_in)
        skips = [h]
        for blk in unet.down_blocks:
            for j, res in enumerate(blk.resnets):
                h = resnet(h, res, temb_act16)
                if blk.attentions is not None:
                    h = transformer(h, blk.attentions[j], ctx16, Lc)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = conv(h, blk.downsamplers[0].conv)
                skips.append(h)
        h = resnet(h, unet.mid_block.resnets[0], temb_act16)
        h = transformer(h, unet.mid_block.attentions[0], ctx16, Lc)
        h = resnet(h, unet.mid_block.resnets[1], temb_act16)
        for blk in unet.up_blocks:
            for j, res in enumerate(blk.resnets):
                h = resnet(torch.cat([h, skips.pop()], dim=-1), res, temb_act16)
                if blk.attentions is not None:
                    h = transformer(h, blk.attentions[j], ctx16, Lc)
            if hasattr(blk, "upsamplers"):
                h = conv(h, blk.upsamplers[0].conv, up=True)
        out = conv(h, unet.conv_out, norm=unet.conv_norm_out)
        return out.permute(0, 3, 1, 2).contiguous()
