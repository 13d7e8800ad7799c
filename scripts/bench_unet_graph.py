"""UNet forward of the SDS step, eager launches vs CUDA-graph replay, CFG pair (batch 2) and one half (batch 1, what a rank of
a multi-GPU job evaluates): python scripts/bench_unet_graph.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avatarcraft_b200.models import sd_unet, sd_native
torch.manual_seed(0)
with torch.device("cuda"):
    unet = sd_unet.UNet2DConditionModel(sd_unet.UNetConfig.sd15()).eval()
t = torch.tensor([417], device="cuda")
for B in (2, 1):
    x = torch.randn(B, 4, 64, 64, device="cuda"); ctx = torch.randn(B, 77, 768, device="cuda")
    for mode in (False, True):
        sd_native.GRAPH = mode
        with torch.no_grad():
            for _ in range(4): unet(x, t, encoder_hidden_states=ctx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): unet(x, t, encoder_hidden_states=ctx)
            e1.record(); e1.synchronize()
        print(f"UNet forward batch {B} {'graph replay' if mode else 'eager launches'}: {e0.elapsed_time(e1) / 10:.2f} ms")
