"""Kernel-time breakdown of the bench's 256x256 stylisation step (16 patches, NeRF side only): python scripts/profile_full_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import synthetic as syn
from avatarcraft_b200.utils.optim import FlatAdam
from avatarcraft_b200.utils.train_utils import native_patch_step

sd = syn.synthetic_state_dict("trained", 43)
net = NeRFNetwork(); net.load_state_dict(sd); net = net.cuda().train()
gt = NeRFNetwork(); gt.load_state_dict(sd); gt = gt.cuda().eval()
for p in gt.parameters(): p.requires_grad_(False)
opt = FlatAdam(net.parameters(), lr=5e-3)
o, d = bench.frame_rays(0)
o, d = o.contiguous().cuda(), d.contiguous().cuda()
G = torch.randn(o.shape[0], 3, device="cuda")
def step():
    native_patch_step(net, gt, opt, o, d, G, batch_size=4096)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); e1.synchronize()
print("pass 2 of 16 patches: %.3f ms" % (e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=26, max_name_column_width=70))
