"""Dense ops of the Stable-Diffusion networks (models/sd_blocks.py), with two explicit execution paths:

  native   CUDA tensor, autograd NOT recording, `NATIVE` on: the hand-written sm_100a kernels of
           libavatarcraft_b200.so (csrc/sd_ops.cu).  Activations are fp32 NHWC (torch channels_last), every GEMM
           operand is fp16 and runs on tcgen05 with fp32 accumulation in TMEM: conv3x3 = fused
           GroupNorm-apply/SiLU/im2col producer + GEMM, linear = LayerNorm/GEGLU/cast producer + GEMM, attention =
           two batched GEMMs around a softmax kernel.  This is the UNet forward of the SDS step
           (`with torch.no_grad()`, models/diffusion.py:121-132).
  autograd torch ops (library kernels), used when a graph is being recorded -- the VAE encoder, whose input gradient
           is the SDS gradient (models/diffusion.py:304-312, :148) -- and on CPU tensors in the unit tests.

There is no silent switch between them: the path is a function of (device, grad mode, NATIVE) only, and the native
path raises if the library is missing."""
import torch
import torch.nn.functional as F

NATIVE = True          # tests flip this to compare the two paths on identical weights


def use_native(x):
    return NATIVE and x.is_cuda and not torch.is_grad_enabled()


def _n():
    from . import sd_native
    return sd_native


def group_norm(x, groups, weight, bias, eps, act):
    if use_native(x):
        return _n().group_norm(x, groups, weight, bias, eps, act)
    y = F.group_norm(x, groups, weight, bias, eps)
    return F.silu(y) if act else y


def layer_norm(x, weight, bias, eps):
    if use_native(x):
        return _n().layer_norm(x, weight, bias, eps)
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def linear(x, weight, bias):
    if use_native(x):
        return _n().linear(x, weight, bias)
    return F.linear(x, weight, bias)


def conv2d(x, weight, bias, stride, padding):
    if use_native(x):
        return _n().conv2d(x, weight, bias, stride, padding)
    return F.conv2d(x, weight, bias, stride=stride, padding=padding)


def add_channel_bias(h, b):
    """h [B,C,H,W] + b [B,C] broadcast over the spatial axes (the time-embedding injection of a resnet block)."""
    return h + b[:, :, None, None]


def geglu(x):
    """diffusers GEGLU: (value, gate) = chunk(x, 2, -1); value * gelu(gate)."""
    if use_native(x):
        return _n().geglu(x)
    v, g = x.chunk(2, dim=-1)
    return v * F.gelu(g)


def attention(q, k, v, heads, scale):
    """q [B,Lq,heads*d], k/v [B,Lk,heads*d] -> softmax(q k^T * scale) v, heads concatenated: [B,Lq,heads*d]."""
    if use_native(q):
        return _n().attention(q, k, v, heads, scale)
    B, Lq, inner = q.shape
    d = inner // heads
    qh = q.reshape(B, Lq, heads, d).transpose(1, 2)
    kh = k.reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, inner)


def to_activation_layout(x):
    """Native path: activations are NHWC in memory (torch channels_last) so that a [B,H,W,C] image is also the
    [B*H*W, C] row-major GEMM operand / token matrix.  Autograd path: unchanged."""
    if use_native(x):
        return x.float().contiguous(memory_format=torch.channels_last)
    return x


def cat_channels(a, b):
    """Skip connection concat along channels."""
    if use_native(a):
        return _n().cat_channels(a, b)
    return torch.cat([a, b], dim=1)
