"""Throughput of the warped (render_can=False) path, BASELINE config 4 shapes: 256x256 rays, 32+32 samples,
synthetic SMPL-shaped body (6890 verts / 13776 faces).  python scripts/bench_warp.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import synthetic as syn
from avatarcraft_b200.utils.render_utils import render_instantnsr_naive
torch.set_grad_enabled(False)
net = NeRFNetwork(); net.load_state_dict(syn.synthetic_state_dict("trained", 43)); net = net.cuda().eval()
net.warp_skip_masked = "--exact-all" not in sys.argv
body = syn.synthetic_body()
o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 256, 256)
o, d = o.cuda(), d.cuda()
def frame():
    return render_instantnsr_naive(net, o, d, 8192, render_can=False, perturb=False, verts=body["world_verts"], faces=body["faces"],
                                   Ts=body["Ts"], num_steps=32, upsample_steps=32, bound=1.6)
for _ in range(2): frame()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): frame()
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": "render_warp animate frame 256x256, 32+32 samples, 13776-triangle posed mesh", "ms_per_frame": ms, "rays_per_sec": 65536 / ms * 1e3}))
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        frame(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
