"""SDS guidance timing on the GPU box: full-size SD-1.5 UNet (random init; weights are not available offline) on the
(uncond, text) latent pair of one 512x512 step -- native sm_100a kernels vs the same network through torch fp32 ops
(the reference's path: diffusers modules in fp32) -- and one complete `mannual_backward` step (VAE encoder with
gradient + UNet + guidance).      python scripts/bench_sds.py [--tiny]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avatarcraft_b200 import _lib
from avatarcraft_b200.models import diffusion, sd_ops, sd_unet, sd_vae

tiny = "--tiny" in sys.argv
torch.manual_seed(0)
cfg = sd_unet.UNetConfig.tiny() if tiny else sd_unet.UNetConfig.sd15()
sd = diffusion.StableDiffusion("cuda", "1.5", unet_config=cfg, vae=sd_vae.AutoencoderKL.tiny() if tiny else None)
emb = sd.get_text_embeds("a bronze statue of a knight")
x = torch.randn(2, 4, 64, 64, device="cuda")
t = torch.tensor([500], device="cuda")


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = {}
with torch.no_grad():
    l0 = _lib.lib().ac_launch_count()
    y = sd.unet(x, t, encoder_hidden_states=emb).sample
    out["native_launches_per_forward"] = int(_lib.lib().ac_launch_count() - l0)
    out["unet_native_ms"] = timed(lambda: sd.unet(x, t, encoder_hidden_states=emb))
    sd_ops.NATIVE = False
    ref = sd.unet(x, t, encoder_hidden_states=emb).sample
    out["unet_torch_fp32_ms"] = timed(lambda: sd.unet(x, t, encoder_hidden_states=emb))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    out["unet_torch_tf32_ms"] = timed(lambda: sd.unet(x, t, encoder_hidden_states=emb))
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    with torch.autocast("cuda", dtype=torch.float16):
        out["unet_torch_autocast_fp16_ms"] = timed(lambda: sd.unet(x, t, encoder_hidden_states=emb))
    sd_ops.NATIVE = True
out["rel_l2_native_vs_torch_fp32"] = float((y.double() - ref.double()).norm() / ref.double().norm())
rgb = torch.rand(1, 3, 256, 256, device="cuda", requires_grad=True)


def step():
    rgb.grad = None
    sd.mannual_backward(emb, rgb, guidance_scale=100)


out["sds_step_native_unet_ms"] = timed(step, n=3, warm=1)
sd_ops.NATIVE = False
out["sds_step_torch_fp32_ms"] = timed(step, n=3, warm=1)
sd_ops.NATIVE = True
out["config"] = "tiny" if tiny else "SD-1.5 UNet 859.5 M params, latents [2,4,64,64], context [2,77,768]; VAE 83.7 M params, 512x512"
out["unet_gflop_per_forward_pair"] = None
print(json.dumps(out))
