"""Kernel-time breakdown of one stylisation step (run on the GPU box): python scripts/profile_train_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import synthetic as syn
from avatarcraft_b200.utils.render_utils import render_instantnsr_naive
from avatarcraft_b200.utils.train_utils import stylize_patch_step

sd = syn.synthetic_state_dict("trained", 43)
net = NeRFNetwork(); net.load_state_dict(sd); net = net.cuda().train()
gt = NeRFNetwork(); gt.load_state_dict(sd); gt = gt.cuda().eval()
for p in gt.parameters(): p.requires_grad_(False)
from avatarcraft_b200.utils.optim import FlatAdam
opt = FlatAdam(net.parameters(), lr=5e-3)
o, d = syn.pinhole_rays(syn.orbit_pose(30.0), 256, 256)
o, d = o.reshape(256, 256, 3)[::4, ::4].reshape(-1, 3).cuda(), d.reshape(256, 256, 3)[::4, ::4].reshape(-1, 3).cuda()
G = torch.randn(o.shape[0], 3, device="cuda")
def step():
    with torch.no_grad():
        render_instantnsr_naive(net, o, d, rays_per_batch=4096, render_can=True, perturb=True)
    stylize_patch_step(net, gt, opt, o, d, G, batch_size=4096)
for _ in range(3): step()
torch.cuda.synchronize()
if "--plain" in sys.argv:          # under ncu: bracket ONE patch step with marker launches so the launch list can be cut
    stylize_patch_step(net, gt, opt, o, d, G, batch_size=4096)
    torch.cuda.synchronize()
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
