"""Time the render kernel of one library build (AC_LIB_PATH) on the bench frame and fingerprint its output.
Run once per variant, on the GPU box:  AC_LIB_PATH=... python scripts/variant_bench.py [tag]
Prints: tag, ms/frame (median of 7, L2 flushed), sha1 of the rgb/depth bits (gather restructurings must not change it)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from avatarcraft_b200.models.instant_nsr import NeRFNetwork
from avatarcraft_b200.utils import synthetic as syn

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("AC_LIB_PATH", "base"))
net = NeRFNetwork(); net.load_state_dict(syn.synthetic_state_dict("trained", 43)); net = net.cuda().eval()
o, d = bench.frame_rays(0); o, d = o.cuda(), d.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.set_grad_enabled(False)
ts = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = net.run(o[None], d[None], 64, 1.6, 64, None, 1.0, 0.0, per_sample_outputs=False)
    e1.record(); e1.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1))
h = hashlib.sha1(out[3].cpu().numpy().tobytes() + out[0].cpu().numpy().tobytes()).hexdigest()[:12]
print(f"VARIANT {tag}: {sorted(ts)[len(ts)//2]:.3f} ms/frame  min {min(ts):.3f}  sha1 {h}")
