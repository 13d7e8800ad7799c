"""CPU suite for the SDS guidance row.  The reference delegates this row to diffusers / HF weights
that are not in /root/reference nor in this image: parity is UNPINNED.  What is pinned here: the closed-form pieces of
the wrapper (noise schedule, add_noise, classifier-free guidance, w(t), clamp, the manual backward) and the
architecture bookkeeping (parameter names / shapes of the published SD-1.5 and SD-2-depth UNets and the VAE)."""
import math

import pytest
import torch

from avatarcraft_b200.models import diffusion, sd_unet, sd_vae


def test_schedule_matches_closed_form():
    ac = diffusion.scaled_linear_alphas_cumprod()
    assert ac.shape == (1000,)
    b0, b1 = 0.00085 ** 0.5, 0.012 ** 0.5
    prod = 1.0
    for i in range(1000):
        prod *= 1.0 - (b0 + (b1 - b0) * i / 999.0) ** 2
        if i in (0, 499, 999):
            assert abs(float(ac[i]) - prod) < 2e-6 * max(prod, 1e-3) + 1e-7
    assert abs(float(ac[0]) - 0.99915) < 1e-6 and float(ac[999]) < 0.005     # published SD values: 0.99915 ... 0.00466


def test_add_noise_and_unet_shapes_tiny():
    torch.manual_seed(0)
    cfg = sd_unet.UNetConfig.tiny()
    unet = sd_unet.UNet2DConditionModel(cfg).eval()
    x = torch.randn(2, 4, 16, 16)
    ctx = torch.randn(2, 7, cfg.cross_attention_dim)
    with torch.no_grad():
        y = unet(x, torch.tensor([500]), encoder_hidden_states=ctx)
    assert y.sample.shape == (2, 4, 16, 16) and y["sample"] is y.sample and torch.isfinite(y.sample).all()
    s = diffusion.NoiseSchedule()
    t = torch.tensor([123])
    z, n = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    a = float(s.alphas_cumprod[123])
    assert torch.allclose(s.add_noise(z, n, t), math.sqrt(a) * z + math.sqrt(1 - a) * n, atol=1e-6)


def test_published_architectures_have_the_published_parameter_counts():
    """SD-1.5 UNet: 859 520 964 parameters; SD VAE: 83 653 863 (diffusers model cards).  Built on the meta device."""
    with torch.device("meta"):
        u15 = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.sd15())
        vae = sd_vae.AutoencoderKL()
        u2d = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.sd2_depth())
    assert sum(p.numel() for p in u15.parameters()) == 859_520_964
    assert sum(p.numel() for p in vae.parameters()) == 83_653_863
    assert sum(p.numel() for p in u2d.parameters()) == 865_910_724 + 320 * 9        # SD-2 base UNet + one extra input channel
    sd = u15.state_dict()
    for k, shape in {"conv_in.weight": (320, 4, 3, 3), "time_embedding.linear_1.weight": (1280, 320),
                     "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight": (320, 768),
                     "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight": (2560, 320),
                     "down_blocks.2.downsamplers.0.conv.weight": (1280, 1280, 3, 3),
                     "mid_block.attentions.0.proj_in.weight": (1280, 1280, 1, 1),
                     "up_blocks.1.resnets.2.conv1.weight": (1280, 1920, 3, 3),
                     "up_blocks.3.resnets.0.conv_shortcut.weight": (320, 960, 1, 1),
                     "up_blocks.2.upsamplers.0.conv.weight": (640, 640, 3, 3),
                     "conv_norm_out.weight": (320,), "conv_out.weight": (4, 320, 3, 3)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "down_blocks.3.attentions.0.norm.weight" not in sd and "up_blocks.0.attentions.0.norm.weight" not in sd
    vs = vae.state_dict()
    assert tuple(vs["encoder.mid_block.attentions.0.to_q.weight"].shape) == (512, 512)
    assert tuple(vs["encoder.conv_out.weight"].shape) == (8, 512, 3, 3) and tuple(vs["quant_conv.weight"].shape) == (8, 8, 1, 1)
    assert tuple(vs["decoder.up_blocks.3.resnets.0.conv_shortcut.weight"].shape) == (128, 256, 1, 1)


def test_vae_loads_0_16_attention_names():
    vae = sd_vae.AutoencoderKL.tiny()
    sd = {k.replace("to_q", "query").replace("to_k", "key").replace("to_v", "value").replace("to_out.0", "proj_attn"): v
          for k, v in vae.state_dict().items()}
    assert any(".query." in k for k in sd)
    sd_vae.AutoencoderKL.tiny().load_state_dict(sd)


class _StubUNet(torch.nn.Module):
    """eps = 0.1 * input + mean(context): lets the SDS arithmetic be written in closed form."""
    in_channels = 4

    def forward(self, x, t, encoder_hidden_states):
        return sd_unet.UNetOutput(0.1 * x + encoder_hidden_states.mean(dim=(1, 2)).reshape(-1, 1, 1, 1))


def test_sds_gradient_arithmetic_and_manual_backward():
    torch.manual_seed(1)
    sd = diffusion.StableDiffusion("cpu", "1.5", unet=_StubUNet(), vae=sd_vae.AutoencoderKL.tiny(),
                                   text_encoder=diffusion.HashTextEncoder(32))
    emb = torch.stack([torch.full((77, 32), 0.2), torch.full((77, 32), 0.5)])         # (uncond, text)
    lat, noise, t = torch.randn(1, 4, 8, 8) * 0.1, torch.randn(1, 4, 8, 8) * 0.01, torch.tensor([300])
    g = sd.sds_latent_gradient(lat, emb, t, noise, guidance_scale=100)
    a = float(sd.alphas[300])
    noisy = math.sqrt(a) * lat + math.sqrt(1 - a) * noise
    eu, et = 0.1 * noisy + 0.2, 0.1 * noisy + 0.5
    want = ((1 - a) * (eu + 100 * (et - eu) - noise)).clamp(-1, 1)
    assert torch.allclose(g, want, atol=1e-5)
    assert float(g.max()) == 1.0                                                       # the clamp is active at scale 100
    # manual backward: d/d(rgb) of <latents, grad> through resize + VAE encoder, with the step's own random draws
    rgb = torch.rand(1, 3, 32, 32, requires_grad=True)
    torch.manual_seed(7)
    sd.mannual_backward(emb, rgb, guidance_scale=100)
    got = rgb.grad.clone()
    torch.manual_seed(7)
    rgb2 = rgb.detach().clone().requires_grad_(True)
    x512 = torch.nn.functional.interpolate(rgb2, (512, 512), mode="bilinear", align_corners=False)
    t2 = torch.randint(sd.min_step, sd.max_step + 1, [1])
    lat2 = sd.encode_imgs(x512)
    noise2 = torch.randn_like(lat2)
    g2 = sd.sds_latent_gradient(lat2.detach(), emb, t2, noise2, 100)
    (lat2 * g2).sum().backward()
    assert torch.allclose(got, rgb2.grad, atol=1e-7) and float(got.abs().max()) > 0
    assert torch.equal(sd.calc_grad(emb, torch.rand(1, 3, 32, 32, requires_grad=True)).isfinite().all(), torch.tensor(True))


def test_text_embeds_are_uncond_then_text():
    sd = diffusion.StableDiffusion("cpu", "1.5", unet=_StubUNet(), vae=sd_vae.AutoencoderKL.tiny(),
                                   text_encoder=diffusion.HashTextEncoder(32))
    e = sd.get_text_embeds("a photo of a knight")
    assert e.shape == (2, 77, 32)
    assert torch.equal(e[:1], sd.get_text_embeds("something else")[:1]) and not torch.equal(e[0], e[1])


def test_fp16_weight_cache_follows_the_parameter_object():
    """The native path caches fp16 GEMM operands per parameter: keyed by weak reference (a new parameter that reuses a
    dead one's id() / address must not see its copy), refreshed on in-place updates, conv kernels permuted tap-major."""
    import gc
    from avatarcraft_b200.models import sd_native
    n0 = len(sd_native._W16)
    p = torch.nn.Parameter(torch.randn(8, 16))
    a = sd_native._w16(p, "linear")
    assert a.dtype == torch.float16 and sd_native._w16(p, "linear") is a and len(sd_native._W16) == n0 + 1
    with torch.no_grad():
        p.add_(1.0)
    b = sd_native._w16(p, "linear")
    assert b is not a and torch.allclose(b.float(), p.detach(), atol=1e-2)
    w = torch.nn.Parameter(torch.randn(5, 4, 3, 3))
    c = sd_native._w16(w, "conv3")                                   # [N, ky, kx, Cp] with Cp = 8, zero padded
    assert c.shape == (5, 72) and torch.equal(c.reshape(5, 3, 3, 8)[..., 4:], torch.zeros(5, 3, 3, 4, dtype=torch.float16))
    assert torch.allclose(c.reshape(5, 3, 3, 8)[..., :4].float(), w.detach().permute(0, 2, 3, 1), atol=1e-2)
    del p, w, a, b, c
    gc.collect()
    assert len(sd_native._W16) == n0
