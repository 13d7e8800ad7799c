"""GPU suite for the SDS guidance row.  Parity is UNPINNED for this row (diffusers / HF weights are
third-party and absent); these tests pin the native sm_100a kernels (csrc/sd_ops.cu) against the SAME network
evaluated with torch fp32 ops on identical random weights.  Tolerances: GEMM operands are fp16 (2^-11 relative per
operand), accumulation fp32."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from avatarcraft_b200 import _lib
from avatarcraft_b200.models import diffusion, sd_native, sd_ops, sd_unet, sd_vae

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 200, 136), (77, 40, 320), (4096, 320, 2880), (2, 1280, 320), (513, 4, 72)])
def test_tcgen05_gemm_against_fp64(shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda().half()
    W = torch.randn(N, K, generator=g).cuda().half()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    gb = torch.randn((M + 63) // 64, N, generator=g).cuda()
    out = sd_native.gemm(A, W, M, N, K, bias=bias, group_bias=gb, rows_per_group=64, residual=res)
    ref = A.double() @ W.double().t() + bias.double() + res.double() + gb.double().repeat_interleave(64, 0)[:M]
    assert float((out.double() - ref).abs().max()) < 2e-3 * (K ** 0.5), shape      # fp32 accumulation of exact fp16 products
    out16 = sd_native.gemm(A, W, M, N, K, out_f16=True)
    assert rel(out16.float(), (A.double() @ W.double().t()).float()) < 2e-3


def test_batched_strided_gemm_is_per_head_attention():
    B, L, Lk, heads, d = 2, 200, 77, 4, 40
    inner = heads * d
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B * L, inner, generator=g).cuda().half()
    k = torch.randn(B * Lk, inner, generator=g).cuda().half()
    Lp = 80
    scores = torch.full((B, heads, L, Lp), float("nan"), device="cuda")
    sd_native.gemm(q, k, L, Lk, d, out=scores, lda=inner, ldw=inner, ldc=Lp, batch=(B, heads), sA=(L * inner, d), sW=(Lk * inner, d),
                   sC=(heads * L * Lp, L * Lp))
    ref = torch.einsum("blhd,bmhd->bhlm", q.float().reshape(B, L, heads, d), k.float().reshape(B, Lk, heads, d))
    assert torch.isnan(scores[..., Lk:]).all()                                   # padding columns untouched
    assert float((scores[..., :Lk] - ref).abs().max()) < 2e-2


@pytest.mark.parametrize("cfg", [(2, 4, 200, 77, 40), (1, 2, 300, 300, 80), (2, 3, 128, 1000, 64), (1, 2, 130, 129, 16), (1, 1, 64, 4096, 40)])
def test_fused_attention_against_torch(cfg):
    B, heads, Lq, Lk, d = cfg
    inner = heads * d
    g = torch.Generator().manual_seed(Lq + Lk + d)
    q = torch.randn(B * Lq, inner, generator=g).cuda().half()
    k = torch.randn(B * Lk, inner, generator=g).cuda().half()
    v = torch.randn(B, Lk, inner, generator=g).cuda().half()
    Lp = (Lk + 7) // 8 * 8
    vt = torch.full((B, inner, Lp), float("nan"), device="cuda", dtype=torch.float16)       # padding must never be read as data
    vt[:, :, :Lk] = v.transpose(1, 2)
    out = torch.empty(B * Lq, inner, device="cuda", dtype=torch.float16)
    scale = d ** -0.5
    _lib.check(_lib.lib().ac_sd_flash_attention_f16(sd_native._p(q), sd_native._p(k), sd_native._p(vt), sd_native._p(out), B, heads, Lq, Lk, d,
                                                    inner, inner, Lp, inner, scale, _lib.stream_ptr()), "flash")
    ref = sd_ops.attention(q.float().reshape(B, Lq, inner), k.float().reshape(B, Lk, inner), v.float(), heads, scale).reshape(B * Lq, inner)
    assert torch.isfinite(out).all()
    assert float((out.float() - ref).abs().max()) < 6e-3, float((out.float() - ref).abs().max())      # fp16 P and output


@pytest.mark.parametrize("cfg", [(2, 64, 64, 64, 96), (2, 32, 32, 128, 64), (2, 16, 16, 192, 40), (2, 8, 8, 64, 130), (3, 8, 8, 64, 16), (2, 32, 32, 64, 24)])
def test_implicit_gemm_conv3x3_against_torch(cfg):
    """3x3 convolution with the nine shifted activation windows fetched by TMA (zero padding = out-of-bounds fill)."""
    B, H, W, C, N = cfg
    g = torch.Generator().manual_seed(H * W + C + N)
    x = torch.randn(B, H, W, C, generator=g).cuda()
    conv = sd_unet.Conv2d(C, N, 3, padding=1).cuda()
    gb = torch.randn(B, N, generator=g).cuda()
    res = torch.randn(B, H, W, N, generator=g).cuda()
    assert sd_native.IMPLICIT_CONV and sd_native._tile_ok(H, W)
    y = sd_native.conv(x, conv, group_bias=gb, residual=res)
    sd_native.IMPLICIT_CONV = False
    try:
        y_im2col = sd_native.conv(x, conv, group_bias=gb, residual=res)          # the im2col + GEMM path on the same inputs
    finally:
        sd_native.IMPLICIT_CONV = True
    assert float((y - y_im2col).abs().max()) < 1e-3
    x16 = x.half().float()
    ref = F.conv2d(x16.permute(0, 3, 1, 2), conv.weight.half().float(), conv.bias, padding=1).permute(0, 2, 3, 1) + gb[:, None, None, :] + res
    assert float((y - ref).abs().max()) < 2e-3 * (9 * C) ** 0.5, float((y - ref).abs().max())


def test_producers_against_torch():
    g = torch.Generator().manual_seed(4)
    B, H, W, C, G = 2, 12, 10, 64, 8
    x = (torch.randn(B, H, W, C, generator=g) * 2 + 0.5).cuda()
    norm = sd_unet.GroupNormAct(G, C, 1e-5, act=True).cuda()
    norm.weight.data.normal_(1.0, 0.2, generator=None); norm.bias.data.normal_(0.0, 0.2)
    ref_n = F.silu(F.group_norm(x.permute(0, 3, 1, 2), G, norm.weight, norm.bias, 1e-5))          # NCHW
    st = sd_native.gn_stats(x, G, 1e-5)
    xg = x.reshape(B, H * W, G, C // G).permute(0, 2, 1, 3).reshape(B, G, -1)
    assert torch.allclose(st[..., 0], xg.mean(-1), atol=1e-5) and torch.allclose(st[..., 1], (xg.var(-1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-4)
    # 3x3 / stride 1 / pad 1 im2col with fused GroupNorm+SiLU == unfold of the normalised image
    A, Ho, Wo = sd_native.im2col(x, 3, 1, 1, norm=norm)
    cols = F.unfold(ref_n, 3, padding=1).reshape(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)   # (tap, c) order
    assert (Ho, Wo) == (H, W) and float((A.float() - cols).abs().max()) < 4e-3
    # stride 2 and nearest x2 up-sampling, no norm
    A2, Ho2, Wo2 = sd_native.im2col(x, 3, 2, 1, Ho=H // 2, Wo=W // 2)
    cols2 = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1, stride=2).reshape(B, C, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * C)
    assert float((A2.float() - cols2).abs().max()) < 4e-3
    A3, Ho3, Wo3 = sd_native.im2col(x, 3, 1, 1, up=True)
    xu = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    cols3 = F.unfold(xu, 3, padding=1).reshape(B, C, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * C)
    assert (Ho3, Wo3) == (2 * H, 2 * W) and float((A3.float() - cols3).abs().max()) < 4e-3
    # channel counts that are not a multiple of 8 are zero padded (conv_in: 4 or 5 channels)
    x5 = torch.randn(1, 6, 6, 5, generator=g).cuda()
    A5, _, _ = sd_native.im2col(x5, 3, 1, 1)
    c5 = F.unfold(x5.permute(0, 3, 1, 2), 3, padding=1).reshape(1, 5, 9, 36).permute(0, 3, 2, 1)
    assert A5.shape == (36, 72) and float((A5.float().reshape(1, 36, 9, 8)[..., :5] - c5).abs().max()) < 2e-3
    assert float(A5.float().reshape(1, 36, 9, 8)[..., 5:].abs().max()) == 0.0
    # LayerNorm, GEGLU, softmax
    t = torch.randn(300, 320, generator=g).cuda()
    ln = torch.nn.LayerNorm(320).cuda()
    assert float((sd_native.layer_norm16(t, ln).float() - ln(t)).abs().max()) < 4e-3
    u = torch.randn(300, 256, generator=g).cuda()
    g16 = torch.empty(300, 128, device="cuda", dtype=torch.float16)
    _lib.check(_lib.lib().ac_sd_geglu_f16(sd_native._p(u), 300, 128, sd_native._p(g16), _lib.stream_ptr()), "geglu")
    assert float((g16.float() - sd_ops.geglu(u)).abs().max()) < 4e-3
    s = torch.randn(64, 80, generator=g).cuda() * 3
    p16 = torch.empty(64, 80, device="cuda", dtype=torch.float16)
    _lib.check(_lib.lib().ac_sd_softmax_f16(sd_native._p(s), 64, 77, 80, 80, 0.3, sd_native._p(p16), _lib.stream_ptr()), "softmax")
    assert float((p16[:, :77].float() - torch.softmax(s[:, :77] * 0.3, -1)).abs().max()) < 1e-3 and float(p16[:, 77:].abs().max()) == 0.0
    s2 = torch.randn(40, 1024, generator=g).cuda() * 4                                  # 16-byte path
    p2 = torch.empty(40, 1024, device="cuda", dtype=torch.float16)
    _lib.check(_lib.lib().ac_sd_softmax_f16(sd_native._p(s2), 40, 1024, 1024, 1024, 0.158, sd_native._p(p2), _lib.stream_ptr()), "softmax")
    assert float((p2.float() - torch.softmax(s2 * 0.158, -1)).abs().max()) < 1e-3 and abs(float(p2.float().sum(-1).mean()) - 1.0) < 2e-3


@pytest.mark.parametrize("linear_proj", [False, True])
def test_native_unet_matches_torch_ops_on_identical_weights(linear_proj):
    torch.manual_seed(0)
    cfg = sd_unet.UNetConfig.tiny()
    cfg.use_linear_projection = linear_proj
    unet = sd_unet.UNet2DConditionModel(cfg).cuda().eval()
    x = torch.randn(2, 4, 32, 32, device="cuda")
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, device="cuda")
    t = torch.tensor([417], device="cuda")
    before = _lib.lib().ac_launch_count()
    with torch.no_grad():
        y = unet(x, t, encoder_hidden_states=ctx).sample
        launches = _lib.lib().ac_launch_count() - before
        sd_ops.NATIVE = False
        try:
            ref = unet(x, t, encoder_hidden_states=ctx).sample
        finally:
            sd_ops.NATIVE = True
    assert launches > 100, "the no-grad CUDA forward must run on libavatarcraft_b200.so"
    assert y.shape == ref.shape == (2, 4, 32, 32)
    assert rel(y, ref) < 1e-2, rel(y, ref)


def test_native_unet_graph_replay_equals_eager_launches():
    """After two eager calls per input signature the native forward is captured into a CUDA graph and replayed: same kernels,
    so replays on NEW inputs (sample, timestep, context all changed) must agree with the eager launches to the forward's own
    run-to-run noise (two eager runs of this net differ by 0 .. 7e-4 relative L2: the split-K GEMMs reduce with atomics and a
    last-bit difference flips fp16 roundings downstream; a stale input would show up as O(0.1)), the launch counter must keep
    counting, and a weight change must drop the graph."""
    from avatarcraft_b200.models import sd_native
    torch.manual_seed(0)
    unet = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.tiny()).cuda().eval()
    assert sd_native.GRAPH
    gen = torch.Generator(device="cuda").manual_seed(5)
    outs, ins = [], []
    with torch.no_grad():
        for i in range(5):
            x = torch.randn(2, 4, 32, 32, device="cuda", generator=gen)
            ctx = torch.randn(2, 77, unet.config.cross_attention_dim, device="cuda", generator=gen)
            t = torch.tensor([100 + 150 * i], device="cuda")
            before = _lib.lib().ac_launch_count()
            outs.append(unet(x, t, encoder_hidden_states=ctx).sample.clone())
            assert _lib.lib().ac_launch_count() - before > 100
            ins.append((x, t, ctx))
        graphs = [v for v in unet._native_graphs.values() if isinstance(v, sd_native._UNetGraph)]
        assert len(graphs) == 1                                   # calls 3..5 were replays
        sd_native.GRAPH = False
        try:
            for (x, t, ctx), y in zip(ins, outs):
                ref = unet(x, t, encoder_hidden_states=ctx).sample
                assert rel(y, ref) < 3e-3, rel(y, ref)
        finally:
            sd_native.GRAPH = True
        unet.conv_out.bias.add_(1.0)                              # in-place weight change: the stale graph must not be used
        x, t, ctx = ins[-1]
        y = unet(x, t, encoder_hidden_states=ctx).sample
        assert rel(y, outs[-1] + 1.0) < 3e-3
        assert not any(isinstance(v, sd_native._UNetGraph) for v in unet._native_graphs.values())


def test_sds_step_runs_end_to_end_with_native_unet():
    torch.manual_seed(0)
    sd = diffusion.StableDiffusion("cuda", "1.5", unet_config=sd_unet.UNetConfig.tiny(), vae=sd_vae.AutoencoderKL.tiny())
    emb = sd.get_text_embeds("a bronze statue")
    rgb = torch.rand(1, 3, 64, 64, device="cuda", requires_grad=True)
    before = _lib.lib().ac_launch_count()
    sd.mannual_backward(emb, rgb, guidance_scale=100)
    assert _lib.lib().ac_launch_count() - before > 100
    assert rgb.grad is not None and torch.isfinite(rgb.grad).all() and float(rgb.grad.abs().max()) > 0


def _vae_native_vs_torch(vae, size, seed):
    """moments and d(loss)/d(image) of the VAE encoder: native kernels (both directions) vs torch fp32 autograd, same weights."""
    from avatarcraft_b200.models import sd_vae_native
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(1, 3, size, size, generator=g) * 2 - 1).cuda()
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False          # a true fp32 reference
    try:
        xr = x.clone().requires_grad_(True)
        mom_r = vae.quant_conv(vae.encoder(xr))
        R = torch.randn(mom_r.shape, generator=g).cuda()
        (mom_r * R).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    before = _lib.lib().ac_launch_count()
    mom_n, backward = sd_vae_native.encode_moments(vae, x)
    gx = backward(R * 8.0) / 8.0                                   # the caller's power-of-two operand scale
    launches = _lib.lib().ac_launch_count() - before
    return rel(mom_n, mom_r.detach()), rel(gx, xr.grad), launches


def test_native_vae_encoder_forward_backward_tiny():
    torch.manual_seed(1)
    vae = sd_vae.AutoencoderKL.tiny().cuda().eval()
    for p in vae.parameters():
        p.requires_grad_(False)
    e_f, e_b, launches = _vae_native_vs_torch(vae, 64, 3)
    print(f"native VAE encoder (tiny) vs torch fp32: moments rel L2 {e_f:.2e}, image gradient rel L2 {e_b:.2e}, {launches} launches")
    assert launches > 50 and e_f < 5e-3 and e_b < 2e-2


def test_native_vae_encoder_forward_backward_sd_size():
    """The VAE of Stable Diffusion 1.5 / 2.x at the SDS step's size (512x512 -> [1,8,64,64]), random weights: both directions
    of the native path against torch fp32 (cuDNN / cuBLAS, TF32 off) autograd on identical weights."""
    torch.manual_seed(2)
    vae = sd_vae.AutoencoderKL().cuda().eval()
    for p in vae.parameters():
        p.requires_grad_(False)
    e_f, e_b, launches = _vae_native_vs_torch(vae, 512, 4)
    print(f"native VAE encoder (SD size, 512x512) vs torch fp32: moments rel L2 {e_f:.2e}, image gradient rel L2 {e_b:.2e}, {launches} launches")
    assert e_f < 5e-3 and e_b < 2e-2


def test_sds_pixel_gradient_native_vae_equals_autograd_vae():
    """StableDiffusion.pixel_gradient with the native VAE (default) vs the torch-autograd VAE on the same seed: same t, noise and
    posterior sample, so the two image gradients must agree to operand rounding."""
    torch.manual_seed(0)
    sd = diffusion.StableDiffusion("cuda", "1.5", unet_config=sd_unet.UNetConfig.tiny(), vae=sd_vae.AutoencoderKL.tiny())
    emb = sd.get_text_embeds("a bronze statue")
    rgb = torch.rand(64 * 64, 3, device="cuda")
    g_native = sd.pixel_gradient(emb, rgb, 64, 64, 100.0, seed=5)
    sd.native_vae = False
    g_torch = sd.pixel_gradient(emb, rgb, 64, 64, 100.0, seed=5)
    assert torch.isfinite(g_native).all() and float(g_torch.abs().max()) > 0
    e = rel(g_native, g_torch)
    print(f"SDS pixel gradient, native VAE vs torch VAE: rel L2 {e:.2e}")
    assert e < 5e-2           # the clamp(-1, 1) of the latent gradient sits between the two VAE passes: a few latents flip sides


def test_native_unet_sd15_size_against_torch_fp32():
    """The full SD-1.5-shaped UNet (859.5 M parameters, random init) on the (uncond, text) pair of the SDS step: native tcgen05
    forward vs torch fp32 ops on identical weights (rel L2 <= 2e-3: fp16 operands, fp32 accumulation, ~60 GEMM layers)."""
    torch.manual_seed(0)
    with torch.device("cuda"):
        unet = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.sd15()).eval()
    x = torch.randn(2, 4, 64, 64, device="cuda")
    ctx = torch.randn(2, 77, 768, device="cuda")
    t = torch.tensor([417], device="cuda")
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    with torch.no_grad():
        y = unet(x, t, encoder_hidden_states=ctx).sample
        sd_ops.NATIVE = False
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            ref = unet(x, t, encoder_hidden_states=ctx).sample
        finally:
            sd_ops.NATIVE = True
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    e = rel(y, ref)
    print(f"native UNet (SD-1.5 size) vs torch fp32: rel L2 {e:.2e}")
    assert e < 2e-3
