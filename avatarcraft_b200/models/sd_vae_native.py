"""Native (sm_100a) forward AND backward of the Stable-Diffusion VAE encoder for the SDS step.

The reference encodes the rendered image WITH gradient (models/diffusion.py:304-312, `posterior.sample() * 0.18215`) and
back-propagates the latent-space SDS gradient to the pixels (`latents.backward(gradient=grad)`, :148).  The VAE weights are
frozen, so the backward consists of DATA gradients only.  Here both directions run on the kernels of csrc/sd_ops.cu:

  forward   conv_in -> 4 encoder blocks (2 resnets each, 3 stride-2 down-samplers) -> mid block (resnet, single-head
            attention over 64x64 tokens, resnet) -> GroupNorm+SiLU -> conv_out -> quant_conv: the same producers and
            TMA / tcgen05 GEMMs as the UNet (sd_native.py); 3x3 convolutions are implicit GEMMs at every resolution.
  backward  walks the tape in reverse: the input gradient of a 3x3 stride-1 convolution is the same implicit-GEMM kernel with
            the spatially flipped, in/out-transposed weights; GroupNorm(+SiLU) backward is a two-pass kernel
            (ac_sd_group_norm_backward); the stride-2 convolutions use a gathered operand (ac_sd_conv_s2_dgrad_operand_f16);
            the attention backward is five GEMMs around ac_sd_softmax_backward_f16.

Gradients travel as fp32 between ops and as fp16 GEMM operands; the caller scales the latent gradient by a power of two so the
fp16 operands keep their precision (everything is linear in it) and divides the image gradient by the same factor.
Activations are NHWC fp32.  No fallback: CUDA tensors and libavatarcraft_b200.so are required."""
import torch

from .. import _lib
from . import sd_native as sn
from .sd_native import _check, _p, _w16, cast16, gemm, gn_stats, im2col


def _w16x(param, kind):
    """Extra fp16 weight layouts for the backward, cached by sd_native._w16's mechanism (keyed per parameter and kind)."""
    slot = sn._W16.get(param)
    if slot is None:
        slot = {}
        sn._W16[param] = slot
    hit = slot.get(kind)
    key = (param.data_ptr(), param._version, tuple(param.shape))
    if hit is not None and hit[0] == key:
        return hit[1]
    w = param.detach()
    if kind == "conv3_dgrad":          # [N,C,3,3] -> rows = input channel c, K = (ky', kx', n) with the taps flipped
        N, C = w.shape[0], w.shape[1]
        Np = (N + 7) // 8 * 8
        buf = torch.zeros(C, 3, 3, Np, device=w.device, dtype=torch.float16)
        buf[..., :N] = w.flip(2, 3).permute(1, 2, 3, 0)
        w16 = buf.reshape(C, 9 * Np)
    elif kind == "conv3_s2_dgrad":     # [N,C,3,3] -> rows = c, K = (ky, kx, n), taps NOT flipped (the operand kernel indexes them)
        N, C = w.shape[0], w.shape[1]
        Np = (N + 7) // 8 * 8
        buf = torch.zeros(C, 3, 3, Np, device=w.device, dtype=torch.float16)
        buf[..., :N] = w.permute(1, 2, 3, 0)
        w16 = buf.reshape(C, 9 * Np)
    elif kind == "T":                  # linear [N,K] or 1x1 conv [N,C,1,1] -> [K, N]
        w16 = w.reshape(w.shape[0], w.shape[1]).t().to(torch.float16).contiguous()
        if w16.shape[1] % 8:
            w16 = torch.nn.functional.pad(w16, (0, 8 - w16.shape[1] % 8)).contiguous()
    else:
        raise ValueError(kind)
    if w16.shape[0] % 8:               # output columns of the GEMM: pad rows so tiny channel counts (3, 8) stay legal
        w16 = torch.nn.functional.pad(w16, (0, 0, 0, 8 - w16.shape[0] % 8)).contiguous()
    slot[kind] = (key, w16)
    return w16


def _norm_cast(x, norm, stats):
    """fp16 NHWC of act(GroupNorm(x)) given precomputed stats."""
    B, H, W, C = x.shape
    out = torch.empty(B * H * W, C, device=x.device, dtype=torch.float16)
    _check(_lib.lib().ac_sd_im2col_f16(_p(x), B, H, W, C, 1, 1, 0, 0, H, W, _p(stats), _p(norm.weight.detach()), _p(norm.bias.detach()),
                                       norm.num_groups, int(norm.act), _p(out), _lib.stream_ptr()), "ac_sd_im2col_f16")
    return out


def _conv3(x32, w16, N, norm=None, stats=None, bias=None, residual=None):
    """3x3 stride-1 pad-1 convolution of NHWC fp32 `x32` (optionally through act(GroupNorm(.)) with precomputed stats) with a
    [N, 9*Cp] weight -> fp32 [B,H,W,N].  Implicit GEMM (TMA-shifted windows of the fp16 activation, no im2col buffer) when the
    channel count and the map shape allow, else im2col + GEMM (conv_in's 3 channels, conv_out's 8-channel gradient, tiny test
    models)."""
    B, H, W, C = x32.shape
    res = None if residual is None else residual.reshape(-1, N)
    if C % 64 == 0 and H == W and (W % 128 == 0 or sn._tile_ok(H, W)) and w16.shape[0] >= N:
        act16 = _norm_cast(x32, norm, stats) if norm is not None else cast16(x32.reshape(-1, C))
        out = torch.empty(B, H, W, N, device=x32.device, dtype=torch.float32)
        _check(_lib.lib().ac_sd_conv3x3_f16(_p(act16), _p(w16), _p(bias), None, _p(res), _p(out), B, H, W, C, N, _lib.stream_ptr()),
               "ac_sd_conv3x3_f16")
        return out
    A, _, _ = im2col(x32, 3, 1, 1, norm=norm)
    return gemm(A, w16, B * H * W, N, A.shape[1], bias=bias, residual=res).reshape(B, H, W, N)


def _gn_backward(x, stats, norm, dy, add=None, want32=True, want16=False):
    B, H, W, C = x.shape
    dev = x.device
    ws = torch.empty(2 * B * norm.num_groups, device=dev, dtype=torch.float64)
    d32 = torch.empty_like(x) if want32 else None
    d16 = torch.empty(B * H * W, C, device=dev, dtype=torch.float16) if want16 else None
    _check(_lib.lib().ac_sd_group_norm_backward(_p(x), _p(dy.contiguous()), B, H * W, C, norm.num_groups, _p(stats), _p(norm.weight.detach()),
                                                _p(norm.bias.detach()), int(norm.act), _p(add), _p(d32), _p(d16), _p(ws), _lib.stream_ptr()),
           "ac_sd_group_norm_backward")
    return d32, d16


def _transpose16(a, rows, cols):
    out = torch.empty(cols, rows, device=a.device, dtype=torch.float16)
    _check(_lib.lib().ac_sd_transpose_f16(_p(a), rows, cols, cols, _p(out), rows, _lib.stream_ptr()), "ac_sd_transpose_f16")
    return out


class _Tape:
    """Reverse-mode tape: each forward op appends a closure d_out -> d_in."""

    def __init__(self):
        self.ops = []

    def backward(self, g):
        for op in reversed(self.ops):
            g = op(g)
        return g


def _resnet(x, mod, tape):
    B, H, W, Cin = x.shape
    Cout = mod.conv1.out_channels
    st1 = gn_stats(x, mod.norm1.num_groups, mod.norm1.eps)
    h = _conv3(x, _w16(mod.conv1.weight, "conv3"), Cout, norm=mod.norm1, stats=st1, bias=mod.conv1.bias.detach())
    st2 = gn_stats(h, mod.norm2.num_groups, mod.norm2.eps)
    if mod.conv_shortcut is None:
        sc = x
    else:
        sc = gemm(cast16(x.reshape(-1, Cin)), _w16(mod.conv_shortcut.weight, "conv1"), B * H * W, Cout, Cin,
                  bias=mod.conv_shortcut.bias.detach()).reshape(B, H, W, Cout)
    out = _conv3(h, _w16(mod.conv2.weight, "conv3"), Cout, norm=mod.norm2, stats=st2, bias=mod.conv2.bias.detach(), residual=sc)

    def bwd(d_out):
        d_out = d_out.contiguous()
        d_a2 = _conv3(d_out, _w16x(mod.conv2.weight, "conv3_dgrad"), Cout)
        d_h, _ = _gn_backward(h, st2, mod.norm2, d_a2)
        d_a1 = _conv3(d_h, _w16x(mod.conv1.weight, "conv3_dgrad"), Cin)
        if mod.conv_shortcut is None:
            d_sc = d_out
        else:
            d_sc = gemm(cast16(d_out.reshape(-1, Cout)), _w16x(mod.conv_shortcut.weight, "T"), B * H * W, Cin, Cout).reshape(B, H, W, Cin)
        d_x, _ = _gn_backward(x, st1, mod.norm1, d_a1, add=d_sc)
        return d_x
    tape.ops.append(bwd)
    return out


def _downsample(x, conv, tape):
    """Downsample2D(padding=0): zero pad bottom/right by one, 3x3 stride-2 valid convolution."""
    B, H, W, C = x.shape
    N, Ho, Wo = conv.out_channels, H // 2, W // 2
    A, _, _ = im2col(x, 3, 2, 0, False, Ho, Wo)
    out = gemm(A, _w16(conv.weight, "conv3"), B * Ho * Wo, N, A.shape[1], bias=conv.bias.detach()).reshape(B, Ho, Wo, N)

    def bwd(d_out):
        Np = (N + 7) // 8 * 8
        op = torch.empty(B * H * W, 9 * Np, device=x.device, dtype=torch.float16)
        _check(_lib.lib().ac_sd_conv_s2_dgrad_operand_f16(_p(d_out.contiguous()), B, Ho, Wo, N, H, W, _p(op), _lib.stream_ptr()),
               "ac_sd_conv_s2_dgrad_operand_f16")
        return gemm(op, _w16x(conv.weight, "conv3_s2_dgrad"), B * H * W, C, 9 * Np).reshape(B, H, W, C)
    tape.ops.append(bwd)
    return out


def _attention(x, attn, tape):
    """VaeAttention: x + to_out(softmax(q k^T / sqrt(C)) v), one head over the H*W tokens of each image."""
    B, H, W, C = x.shape
    L = H * W
    st = gn_stats(x, attn.group_norm.num_groups, attn.group_norm.eps)
    xn = _norm_cast(x, attn.group_norm, st)                                            # [B*L, C] fp16
    lin = lambda a16, m, M, **kw: gemm(a16, _w16(m.weight, "linear"), M, m.out_features, m.in_features, bias=m.bias.detach(), **kw)
    q, k, v = (lin(xn, m, B * L, out_f16=True) for m in (attn.to_q, attn.to_k, attn.to_v))
    saved = []
    o = torch.empty(B * L, C, device=x.device, dtype=torch.float16)
    for b in range(B):
        qb, kb, vb = q[b * L:(b + 1) * L], k[b * L:(b + 1) * L], v[b * L:(b + 1) * L]
        scores = gemm(qb, kb, L, L, C)                                                 # [L, L] fp32
        P = torch.empty(L, L, device=x.device, dtype=torch.float16)
        _check(_lib.lib().ac_sd_softmax_f16(_p(scores), L, L, L, L, float(attn.scale), _p(P), _lib.stream_ptr()), "ac_sd_softmax_f16")
        vT = _transpose16(vb, L, C)                                                    # [C, L]
        gemm(P, vT, L, C, L, out_f16=True, out=o[b * L:(b + 1) * L])
        saved.append((qb, kb, vb, P))
    out = lin(o, attn.to_out[0], B * L, residual=x.reshape(-1, C)).reshape(B, H, W, C)

    def bwd(d_out):
        d16 = cast16(d_out.reshape(-1, C))
        d_o = gemm(d16, _w16x(attn.to_out[0].weight, "T"), B * L, C, C, out_f16=True)   # [B*L, C] fp16
        d_xn = torch.empty(B * L, C, device=x.device, dtype=torch.float32)
        for b, (qb, kb, vb, P) in enumerate(saved):
            dob = d_o[b * L:(b + 1) * L]
            dV = gemm(_transpose16(P, L, L), _transpose16(dob, L, C), L, C, L, out_f16=True)            # P^T d_o
            dP = gemm(dob, vb, L, L, C)                                                                 # d_o v^T, fp32
            dS = torch.empty(L, L, device=x.device, dtype=torch.float16)
            _check(_lib.lib().ac_sd_softmax_backward_f16(_p(P), _p(dP), L, L, L, float(attn.scale), _p(dS), _lib.stream_ptr()),
                   "ac_sd_softmax_backward_f16")
            dQ = gemm(dS, _transpose16(kb, L, C), L, C, L, out_f16=True)                                # dS k
            dK = gemm(_transpose16(dS, L, L), _transpose16(qb, L, C), L, C, L, out_f16=True)            # dS^T q
            acc = gemm(dQ, _w16x(attn.to_q.weight, "T"), L, C, C)
            acc = gemm(dK, _w16x(attn.to_k.weight, "T"), L, C, C, residual=acc)
            gemm(dV, _w16x(attn.to_v.weight, "T"), L, C, C, residual=acc, out=d_xn[b * L:(b + 1) * L])
        d_x, _ = _gn_backward(x, st, attn.group_norm, d_xn.reshape(B, H, W, C), add=d_out)
        return d_x
    tape.ops.append(bwd)
    return out


def encode_moments(vae, x):
    """x [B,3,H,W] in [-1,1] (fp32, CUDA) -> (moments [B,2*latent,H/8,W/8] fp32, backward) where backward(d_moments) returns
    d loss / d x [B,3,H,W].  `vae`: models.sd_vae.AutoencoderKL (frozen)."""
    if not x.is_cuda:
        raise RuntimeError("sd_vae_native.encode_moments needs CUDA tensors (no CPU path)")
    enc = vae.encoder
    with torch.no_grad(), torch.cuda.device(x.device):
        tape = _Tape()
        B, _, H, W = x.shape
        xin = x.float().permute(0, 2, 3, 1).contiguous()
        ci = enc.conv_in
        h = _conv3(xin, _w16(ci.weight, "conv3"), ci.out_channels, bias=ci.bias.detach())

        def conv_in_bwd(d):
            return _conv3(d.contiguous(), _w16x(ci.weight, "conv3_dgrad"), 8)[..., :3]      # 3 input channels, weight rows padded to 8
        tape.ops.append(conv_in_bwd)
        for blk in enc.down_blocks:
            for r in blk.resnets:
                h = _resnet(h, r, tape)
            if hasattr(blk, "downsamplers"):
                h = _downsample(h, blk.downsamplers[0].conv, tape)
        mid = enc.mid_block
        h = _resnet(h, mid.resnets[0], tape)
        h = _attention(h, mid.attentions[0], tape)
        h = _resnet(h, mid.resnets[1], tape)
        Bh, Hh, Wh, Ch = h.shape
        no, co, qc = enc.conv_norm_out, enc.conv_out, vae.quant_conv
        st = gn_stats(h, no.num_groups, no.eps)
        Nm = co.out_channels
        m = _conv3(h, _w16(co.weight, "conv3"), Nm, norm=no, stats=st, bias=co.bias.detach())
        h_last = h
        mom = gemm(cast16(m.reshape(-1, Nm)), _w16(qc.weight, "conv1"), Bh * Hh * Wh, Nm, Nm, bias=qc.bias.detach()).reshape(Bh, Hh, Wh, Nm)

        def head_bwd(d):                                    # quant_conv, conv_out, conv_norm_out (+ SiLU)
            dm = gemm(cast16(d.reshape(-1, Nm)), _w16x(qc.weight, "T"), Bh * Hh * Wh, Nm, Nm).reshape(Bh, Hh, Wh, Nm)
            d_a = _conv3(dm, _w16x(co.weight, "conv3_dgrad"), Ch)
            d_h, _ = _gn_backward(h_last, st, no, d_a)
            return d_h
        tape.ops.append(head_bwd)

    def backward(d_moments):
        with torch.no_grad(), torch.cuda.device(x.device):
            d = d_moments.float().permute(0, 2, 3, 1).contiguous()
            return tape.backward(d).permute(0, 3, 1, 2).contiguous()
    return mom.permute(0, 3, 1, 2).contiguous(), backward
