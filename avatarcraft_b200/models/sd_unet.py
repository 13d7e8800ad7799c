"""UNet2DConditionModel of Stable Diffusion (the `unet` of models/diffusion.py:59, diffusers==0.16.1 -- third-party,
absent from /root/reference: restated from the published architecture, **parity unpinned**).

Parameter names / shapes are those of `unet/diffusion_pytorch_model.*`, so a real checkpoint loads with
`load_state_dict(strict=True)`.  `UNetConfig.sd15()` = runwayml/stable-diffusion-v1-5, `UNetConfig.sd2_depth()` =
stabilityai/stable-diffusion-2-depth (the two model ids of models/diffusion.py:45-49); `UNetConfig.tiny()` is a
narrow version for tests.  The forward mirrors UNet2DConditionModel.forward: timestep embedding -> conv_in -> down
blocks (skip stack) -> mid -> up blocks (skip concat) -> GroupNorm/SiLU/conv_out."""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import sd_ops
from .sd_blocks import Conv2d, Downsample2D, GroupNormAct, Linear, ResnetBlock2D, Transformer2DModel, Upsample2D, timestep_embedding


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)      # diffusers' (mis)name: this is the NUMBER of heads per block
    attn_blocks: Tuple[bool, ...] = (True, True, True, False)  # CrossAttnDownBlock2D x3, DownBlock2D
    use_linear_projection: bool = False
    norm_num_groups: int = 32

    @staticmethod
    def sd15():
        return UNetConfig()

    @staticmethod
    def sd2_depth():
        return UNetConfig(in_channels=5, cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20), use_linear_projection=True)

    @staticmethod
    def sd21():
        """Stable Diffusion 2.1 (base and v-pred share it): OpenCLIP-H text width 1024, 64-wide heads, linear projections."""
        return UNetConfig(in_channels=4, cross_attention_dim=1024, attention_head_dim=(5, 10, 20, 20), use_linear_projection=True)

    @staticmethod
    def tiny():
        return UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=48, attention_head_dim=(2, 2, 4, 4), norm_num_groups=8)


class _Block(nn.Module):
    """One down / up block: resnets (+ cross-attention transformers) (+ down/up-sampler)."""

    def __init__(self, res_channels, out_ch, temb_ch, cfg, heads, with_attn, down, up):
        super().__init__()
        g = cfg.norm_num_groups
        self.resnets = nn.ModuleList([ResnetBlock2D(ci, out_ch, temb_ch, g) for ci in res_channels])
        if with_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, out_ch // heads, out_ch, cfg.cross_attention_dim, g,
                                                                cfg.use_linear_projection) for _ in res_channels])
        else:
            self.attentions = None
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(out_ch)])
        if up:
            self.upsamplers = nn.ModuleList([Upsample2D(out_ch)])


class UNetMidBlock(nn.Module):
    def __init__(self, ch, temb_ch, cfg, heads):
        super().__init__()
        g = cfg.norm_num_groups
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, ch, cfg.cross_attention_dim, g, cfg.use_linear_projection)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, g), ResnetBlock2D(ch, ch, temb_ch, g)])


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1, self.linear_2 = Linear(cin, cout), Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = None):
        super().__init__()
        cfg = cfg or UNetConfig.sd15()
        self.config = cfg
        self.in_channels = cfg.in_channels                       # read by produce_latents (models/diffusion.py:270)
        boc = cfg.block_out_channels
        temb_ch = boc[0] * 4
        self.conv_in = Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, ch in enumerate(boc):
            in_ch, out_ch = out_ch, ch
            res_in = [in_ch] + [out_ch] * (cfg.layers_per_block - 1)
            self.down_blocks.append(_Block(res_in, out_ch, temb_ch, cfg, cfg.attention_head_dim[i], cfg.attn_blocks[i],
                                           down=i < len(boc) - 1, up=False))
        self.mid_block = UNetMidBlock(boc[-1], temb_ch, cfg, cfg.attention_head_dim[-1])
        self.up_blocks = nn.ModuleList()
        rev, rev_heads, rev_attn = boc[::-1], cfg.attention_head_dim[::-1], cfg.attn_blocks[::-1]
        out_ch = rev[0]
        for i, ch in enumerate(rev):
            prev_out, out_ch = out_ch, ch
            in_ch = rev[min(i + 1, len(boc) - 1)]
            n = cfg.layers_per_block + 1
            res_in = [(prev_out if j == 0 else out_ch) + (in_ch if j == n - 1 else out_ch) for j in range(n)]
            self.up_blocks.append(_Block(res_in, out_ch, temb_ch, cfg, rev_heads[i], rev_attn[i], down=False, up=i < len(boc) - 1))
        self.conv_norm_out = GroupNormAct(cfg.norm_num_groups, boc[0], 1e-5, act=True)
        self.conv_out = Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states):
        """sample [B,Cin,H,W], timestep scalar / [1] / [B], encoder_hidden_states [B,77,D] -> [B,Cout,H,W]
        (the reference reads `.sample` of diffusers' output object: see UNetOutput)."""
        if sd_ops.use_native(sample):                           # no-grad CUDA call: the hand-written sm_100a kernels
            from . import sd_native
            return UNetOutput(sd_native.unet_forward(self, sample, timestep, encoder_hidden_states))
        B = sample.shape[0]
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(B)
        temb = self.time_embedding(timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype))
        temb_act = F.silu(temb)                                 # every resnet applies SiLU before its projection
        ctx = encoder_hidden_states
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(h, temb_act)
                if blk.attentions is not None:
                    h = blk.attentions[j](h, ctx)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)
        h = self.mid_block.resnets[0](h, temb_act)
        h = self.mid_block.attentions[0](h, ctx)
        h = self.mid_block.resnets[1](h, temb_act)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(sd_ops.cat_channels(h, skips.pop()), temb_act)
                if blk.attentions is not None:
                    h = blk.attentions[j](h, ctx)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        return UNetOutput(self.conv_out(self.conv_norm_out(h)).contiguous())


class UNetOutput:
    """Stand-in for diffusers' UNet2DConditionOutput: `.sample` and `['sample']` (models/diffusion.py:132,283)."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, k):
        if k in (0, "sample"):
            return self.sample
        raise KeyError(k)
