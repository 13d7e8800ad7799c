"""Kernel-time breakdown of one SDS guidance call (StableDiffusion.pixel_gradient at SD-1.5 size: resize, VAE encoder forward,
UNet on the CFG pair, VAE encoder backward) -- run on the GPU box: python scripts/profile_guidance.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avatarcraft_b200.models.diffusion import StableDiffusion

sd = StableDiffusion("cuda", "1.5")
emb = sd.get_text_embeds("a 3D rendering of a knight in bronze armour")
rgb = torch.rand(256 * 256, 3, device="cuda")
for i in range(3):
    sd.pixel_gradient(emb, rgb, 256, 256, 100.0, seed=i)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(5):
    sd.pixel_gradient(emb, rgb, 256, 256, 100.0, seed=10 + i)
ev[1].record(); torch.cuda.synchronize()
print(f"pixel_gradient: {ev[0].elapsed_time(ev[1]) / 5:.2f} ms per call (native VAE = {sd.native_vae})")
from avatarcraft_b200.models import sd_vae_native
x = torch.rand(1, 3, 512, 512, device="cuda") * 2 - 1
for name in ("fwd", "fwd+bwd"):
    for _ in range(2):
        m, bw = sd_vae_native.encode_moments(sd.vae, x)
        if name != "fwd": bw(torch.randn_like(m))
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(5):
        m, bw = sd_vae_native.encode_moments(sd.vae, x)
        if name != "fwd": bw(torch.randn_like(m))
    ev[1].record(); torch.cuda.synchronize()
    print(f"native VAE encoder {name}: {ev[0].elapsed_time(ev[1]) / 5:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sd.pixel_gradient(emb, rgb, 256, 256, 100.0, seed=99)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
