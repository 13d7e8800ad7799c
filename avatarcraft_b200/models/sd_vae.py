"""AutoencoderKL of Stable Diffusion (the `vae` of models/diffusion.py:53; diffusers==0.16.1 -- third-party, absent
from /root/reference: restated from the published architecture, **parity unpinned**).  Parameter names follow
`vae/diffusion_pytorch_model.*` (0.16.1 attention names query/key/value/proj_attn are remapped on load).

The SDS step differentiates THROUGH the encoder (`latents.backward(gradient=grad)`, models/diffusion.py:148), so the
encoder runs on the autograd path of sd_ops; `decode` is only used by prompt_to_img."""
import torch
import torch.nn as nn

from .sd_blocks import Conv2d, Downsample2D, GroupNormAct, ResnetBlock2D, Upsample2D, VaeAttention


class _MidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.attentions = nn.ModuleList([VaeAttention(ch, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, groups, eps=1e-6), ResnetBlock2D(ch, ch, None, groups, eps=1e-6)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, groups, eps=1e-6) for j in range(layers)])
        if down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if hasattr(self, "downsamplers") else x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, None, groups, eps=1e-6) for j in range(layers)])
        if up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if hasattr(self, "upsamplers") else x


class Encoder(nn.Module):
    def __init__(self, cin, latent, boc, layers, groups):
        super().__init__()
        self.conv_in = Conv2d(cin, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList([_EncBlock(boc[max(i - 1, 0)], boc[i], layers, groups, down=i < len(boc) - 1) for i in range(len(boc))])
        self.mid_block = _MidBlock(boc[-1], groups)
        self.conv_norm_out = GroupNormAct(groups, boc[-1], 1e-6, act=True)
        self.conv_out = Conv2d(boc[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        return self.conv_out(self.conv_norm_out(self.mid_block(x)))


class Decoder(nn.Module):
    def __init__(self, latent, cout, boc, layers, groups):
        super().__init__()
        rev = boc[::-1]
        self.conv_in = Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = _MidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList([_DecBlock(rev[max(i - 1, 0)], rev[i], layers + 1, groups, up=i < len(boc) - 1) for i in range(len(boc))])
        self.conv_norm_out = GroupNormAct(groups, rev[-1], 1e-6, act=True)
        self.conv_out = Conv2d(rev[-1], cout, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_norm_out(x))


class DiagonalGaussian:
    """diffusers DiagonalGaussianDistribution: parameters [B, 2C, h, w] = (mean, logvar clamped to [-30, 20])."""

    def __init__(self, params):
        self.mean, logvar = params.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _Encoded:
    def __init__(self, dist):
        self.latent_dist = dist


class _Decoded:
    def __init__(self, sample):
        self.sample = sample


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = Conv2d(latent_channels, latent_channels, 1)

    @staticmethod
    def tiny():
        return AutoencoderKL(block_out_channels=(16, 32, 32, 32), layers_per_block=1, norm_num_groups=8)

    def encode(self, x):
        return _Encoded(DiagonalGaussian(self.quant_conv(self.encoder(x))))

    def decode(self, z):
        return _Decoded(self.decoder(self.post_quant_conv(z)))
