"""Kernel-time breakdown of one native SD-1.5 UNet forward (run on the GPU box): per-kernel totals and a census of the
GEMM shapes with their GPU durations (profiler kernel events zipped with the Python-side shape log, same order)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avatarcraft_b200.models import sd_unet, sd_native

torch.manual_seed(0)
unet = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.sd15()).cuda().eval()
x = torch.randn(2, 4, 64, 64, device="cuda"); t = torch.tensor([500], device="cuda"); emb = torch.randn(2, 77, 768, device="cuda")
shapes = []
orig = sd_native.gemm
def spy(A16, W16, M, N, K, **kw):
    b = kw.get("batch", (1, 1))
    shapes.append((M, N, K, b[0] * b[1]))
    return orig(A16, W16, M, N, K, **kw)
sd_native.gemm = spy
with torch.no_grad():
    for _ in range(2):
        unet(x, t, encoder_hidden_states=emb)
    torch.cuda.synchronize()
    shapes.clear()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        unet(x, t, encoder_hidden_states=emb)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
ev = sorted([e for e in prof.events() if "sd_gemm" in e.name], key=lambda e: e.time_range.start)
assert len(ev) == len(shapes), (len(ev), len(shapes))
census = collections.OrderedDict()
for s, e in zip(shapes, ev):
    c = census.setdefault(s, [0, 0.0]); c[0] += 1; c[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in census.values())
print(f"GEMM total {tot / 1e3:.2f} ms over {len(ev)} launches")
for k, v in sorted(census.items(), key=lambda kv: -kv[1][1])[:28]:
    M, N, K, b = k
    fl = 2.0 * M * N * K * b * v[0]
    print(f"M={M:6d} N={N:5d} K={K:5d} batch={b:3d} calls={v[0]:3d}  {v[1] / 1e3:7.3f} ms  {fl / v[1] / 1e6:8.1f} TFLOP/s  ctas={((M + 127) // 128) * ((N + 127) // 128) * b}")
