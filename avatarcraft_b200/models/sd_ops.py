"""Dense ops of the Stable-Diffusion modules (models/sd_blocks.py) on the AUTOGRAD path, and the switch between the
two execution paths of the UNet:

  native   CUDA tensors, autograd NOT recording, `NATIVE` on: `UNet2DConditionModel.forward` hands the whole forward to
           models/sd_native.unet_forward -- the hand-written sm_100a kernels of csrc/sd_ops.cu (tcgen05 GEMMs fed by
           fused GroupNorm/SiLU/im2col, LayerNorm, GEGLU, softmax producers; fp16 operands, fp32 accumulation).  This
           is the UNet evaluation of the SDS step (`with torch.no_grad()`, models/diffusion.py:121-132).
  autograd the functions below (torch ops, library kernels): used when a graph is being recorded -- the VAE encoder,
           whose input gradient IS the SDS gradient (models/diffusion.py:304-312, :148) -- and for CPU unit tests.

The path is a function of (device, grad mode, NATIVE) only; the native path raises if the library is missing."""
import torch
import torch.nn.functional as F

NATIVE = True          # tests flip this to compare the two paths on identical weights


def use_native(x):
    return NATIVE and x.is_cuda and not torch.is_grad_enabled()


def group_norm(x, groups, weight, bias, eps, act):
    y = F.group_norm(x, groups, weight, bias, eps)
    return F.silu(y) if act else y


def layer_norm(x, weight, bias, eps):
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def linear(x, weight, bias):
    return F.linear(x, weight, bias)


def conv2d(x, weight, bias, stride, padding):
    return F.conv2d(x, weight, bias, stride=stride, padding=padding)


def add_channel_bias(h, b):
    """h [B,C,H,W] + b [B,C] broadcast over the spatial axes (the time-embedding injection of a resnet block)."""
    return h + b[:, :, None, None]


def geglu(x):
    """diffusers GEGLU: (value, gate) = chunk(x, 2, -1); value * gelu(gate)."""
    v, g = x.chunk(2, dim=-1)
    return v * F.gelu(g)


def attention(q, k, v, heads, scale):
    """q [B,Lq,heads*d], k/v [B,Lk,heads*d] -> softmax(q k^T * scale) v, heads concatenated: [B,Lq,heads*d]."""
    B, Lq, inner = q.shape
    d = inner // heads
    qh = q.reshape(B, Lq, heads, d).transpose(1, 2)
    kh = k.reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, inner)


def cat_channels(a, b):
    return torch.cat([a, b], dim=1)
