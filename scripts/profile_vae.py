"""Kernel-time breakdown of the native VAE encoder forward + backward at SD size (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avatarcraft_b200.models import sd_vae, sd_vae_native
torch.manual_seed(0)
vae = sd_vae.AutoencoderKL().cuda().eval()
for p in vae.parameters(): p.requires_grad_(False)
x = torch.rand(1, 3, 512, 512, device="cuda") * 2 - 1
for _ in range(3):
    m, bw = sd_vae_native.encode_moments(vae, x); bw(torch.randn_like(m))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m, bw = sd_vae_native.encode_moments(vae, x); bw(torch.randn_like(m))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
