"""Building blocks of the Stable-Diffusion networks the SDS guidance runs.

The reference instantiates diffusers' `UNet2DConditionModel` and `AutoencoderKL` (models/diffusion.py:53-60;
pip pins diffusers==0.16.1, readme.md:35 -- third-party code that is NOT under /root/reference and not installed in
this image, so this is a from-scratch restatement of the published architecture: **parity unpinned**).  Parameter
names and shapes follow the diffusers checkpoints (`unet/diffusion_pytorch_model.*`, `vae/...`), so real weights
load with `load_state_dict` when they are available.

These modules hold the parameters and define the autograd / torch-op forward (`sd_ops`): the VAE encoder, whose input
gradient IS the SDS gradient, needs a backward.  The no-grad UNet evaluation of the SDS step on a CUDA device does not
run these forwards: `UNet2DConditionModel.forward` hands it to models/sd_native.py, the hand-written sm_100a kernels
(csrc/sd_ops.cu: tcgen05 GEMMs fed by fused GroupNorm/SiLU/im2col, LayerNorm, GEGLU, softmax producers)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sd_ops


class GroupNormAct(nn.GroupNorm):
    """GroupNorm optionally followed by SiLU, fused into one kernel on the native path."""

    def __init__(self, groups, channels, eps, act=False):
        super().__init__(groups, channels, eps=eps, affine=True)
        self.act = act

    def forward(self, x):
        return sd_ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, self.act)


class Linear(nn.Linear):
    def forward(self, x):
        return sd_ops.linear(x, self.weight, self.bias)


class Conv2d(nn.Conv2d):
    def forward(self, x):
        return sd_ops.conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0])


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return sd_ops.layer_norm(x, self.weight, self.bias, self.eps)


def timestep_embedding(t, dim, flip_sin_to_cos=True, freq_shift=0.0, max_period=10000.0):
    """diffusers `Timesteps` / get_timestep_embedding: [B] -> [B, dim] = cat(cos, sin) (flip) of t * 10000^(-i/(half-shift))."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift)
    ang = t.float()[:, None] * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_channels=None, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = GroupNormAct(groups, cin, eps, act=True)
        self.conv1 = Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = GroupNormAct(groups, cout, eps, act=True)
        self.conv2 = Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb_act=None):
        """temb_act = silu(time embedding) [B, temb_channels] (the activation is shared by every block)."""
        h = self.conv1(self.norm1(x))
        if self.time_emb_proj is not None and temb_act is not None:
            h = sd_ops.add_channel_bias(h, self.time_emb_proj(temb_act))
        h = self.conv2(self.norm2(h))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Attention(nn.Module):
    """diffusers CrossAttention / Attention: to_q/to_k/to_v (no bias), to_out.0 (bias)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, qkv_bias=False):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = Linear(query_dim, inner, bias=qkv_bias)
        self.to_k = Linear(context_dim or query_dim, inner, bias=qkv_bias)
        self.to_v = Linear(context_dim or query_dim, inner, bias=qkv_bias)
        self.to_out = nn.ModuleList([Linear(inner, query_dim), nn.Identity()])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        return self.to_out[0](sd_ops.attention(q, k, v, self.heads, self.scale))


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = Linear(dim, inner * 2)

    def forward(self, x):
        return sd_ops.geglu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Identity(), Linear(dim * mult, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, context_dim):
        super().__init__()
        self.norm1 = LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = LayerNorm(dim)
        self.attn2 = Attention(dim, context_dim, heads, dim_head)
        self.norm3 = LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, channels, context_dim, groups=32, use_linear_projection=False):
        super().__init__()
        inner = heads * dim_head
        self.use_linear = use_linear_projection
        self.norm = GroupNormAct(groups, channels, 1e-6, act=False)
        self.proj_in = Linear(channels, inner) if use_linear_projection else Conv2d(channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, context_dim)])
        self.proj_out = Linear(inner, channels) if use_linear_projection else Conv2d(inner, channels, 1)

    def forward(self, x, context):
        B, C, H, W = x.shape
        h = self.norm(x)
        if self.use_linear:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, -1)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        if self.use_linear:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous())
        return h + x


class Downsample2D(nn.Module):
    def __init__(self, channels, padding=1):
        super().__init__()
        self.pad = padding
        self.conv = Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.pad == 0:                              # VAE encoder: asymmetric (0,1,0,1) zero pad, then a valid conv
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class VaeAttention(nn.Module):
    """Single-head spatial self-attention of the VAE mid block.  Parameter names of diffusers >= 0.18
    (group_norm, to_q, to_k, to_v, to_out.0); 0.16.1 checkpoints (query/key/value/proj_attn) are remapped on load."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = GroupNormAct(groups, channels, eps, act=False)
        self.to_q, self.to_k, self.to_v = Linear(channels, channels), Linear(channels, channels), Linear(channels, channels)
        self.to_out = nn.ModuleList([Linear(channels, channels), nn.Identity()])
        self.scale = channels ** -0.5

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = self.to_out[0](sd_ops.attention(self.to_q(h), self.to_k(h), self.to_v(h), 1, self.scale))
        return x + h.reshape(B, H, W, C).permute(0, 3, 1, 2)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for old, new in (("query", "to_q"), ("key", "to_k"), ("value", "to_v"), ("proj_attn", "to_out.0")):
            for leaf in ("weight", "bias"):
                k = f"{prefix}{old}.{leaf}"
                if k in state_dict:
                    state_dict[f"{prefix}{new}.{leaf}"] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
