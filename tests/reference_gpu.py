"""The reference's GPU path on the B200, as a checker and as "the number to beat" (SURVEY.md 8d).

`RefGpuNSR` runs the oracle's restatement of NeRFRenderer.run as ~250 eager torch launches per batch
on CUDA tensors, with the hash encode done by the REFERENCE'S OWN kernel (oracle/_ref/_ref_hash_encoder.so,
unmodified hashencoder.cu built for sm_100a) through the same [L,B,C] -> permute -> [B,L*C] sequence as the
reference's HashEncoder.forward (encoder/hashencoder/hashgrid.py:126-142).  Test infrastructure only.

Run as a script on the GPU box to time it next to the fused kernel:
    python tests/reference_gpu.py  ->  gpurun_out/reference_gpu.json
"""
import importlib.machinery
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.nsr_oracle import OracleNSR  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "_ref_hash_encoder.so")


def ref_hash_module():
    loader = importlib.machinery.ExtensionFileLoader("_ref_hash_encoder", REF_SO)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("_ref_hash_encoder", loader))
    loader.exec_module(mod)
    return mod


class RefGpuNSR(OracleNSR):
    def __init__(self, state_dict):
        super().__init__(state_dict)
        self.ref = ref_hash_module()
        mv = lambda t: t.cuda()
        self.table, self.offsets, self.variance = mv(self.table), mv(self.offsets), mv(self.variance)
        self.sdf_w, self.sdf_b, self.col_w = [mv(t) for t in self.sdf_w], [mv(t) for t in self.sdf_b], [mv(t) for t in self.col_w]
        self.S = float(np.log2(self.per_level_scale))

    def encode(self, x, bound, want_ids=False):
        x01 = ((x + bound) / (2 * bound)).contiguous()
        B, L = x01.shape[0], self.offsets.numel() - 1
        out = torch.empty(L, B, 2, device="cuda")
        dummy = torch.empty(1, device="cuda")
        self.ref.hash_encode_forward(x01, self.table, self.offsets, out, B, 3, 2, L, self.S, self.base_resolution, False, dummy)
        return out.permute(1, 0, 2).reshape(B, L * 2)

    @torch.no_grad()
    def render_frame(self, rays_o, rays_d, num_steps=64, upsample_steps=64, bound=1.6, batch=4096):
        """The reference's render_instantnsr_naive loop shape (utils/render_utils.py:514-600): 4096-ray batches."""
        rgb, depth, wsum = [], [], []
        with torch.device("cuda"):
            for s in range(0, rays_o.shape[0], batch):
                out = self.run(rays_o[s:s + batch], rays_d[s:s + batch], num_steps, bound, upsample_steps)
                rgb.append(out[3].reshape(-1, 3)); depth.append(out[0].reshape(-1)); wsum.append(out[2].reshape(-1))
        return torch.cat(rgb), torch.cat(depth), torch.cat(wsum)


def main():
    from avatarcraft_b200.utils import synthetic as syn
    from tests.util import gpu_model, psnr
    sd = syn.synthetic_state_dict("trained", 43)
    o, d = syn.pinhole_rays(syn.orbit_pose(0.0), 256, 256)
    o, d = o.cuda(), d.cuda()
    ref = RefGpuNSR(sd)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(2):
        rgb_r, _, _ = ref.render_frame(o, d)
    torch.cuda.synchronize()
    ev[0].record()
    reps = 3
    for _ in range(reps):
        rgb_r, dep_r, ws_r = ref.render_frame(o, d)
    ev[1].record(); torch.cuda.synchronize()
    ms_ref = ev[0].elapsed_time(ev[1]) / reps
    net = gpu_model(sd)
    with torch.no_grad():
        for _ in range(3):
            out = net.run(o[None], d[None], 64, 1.6, 64, None, 1.0, 0.0)
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(10):
            out = net.run(o[None], d[None], 64, 1.6, 64, None, 1.0, 0.0)
        ev[1].record(); torch.cuda.synchronize()
    ms_mine = ev[0].elapsed_time(ev[1]) / 10
    res = {"workload": "C2 256x256 rays, 64+64 samples, bound 1.6, white bg, trained-like synthetic checkpoint",
           "reference_gpu_eager": {"ms_per_frame": ms_ref, "rays_per_s": 65536 / ms_ref * 1e3,
                                   "what": "reference hashencoder.cu (unmodified, sm_100a) + eager torch fp32 restatement of run(), 4096-ray batches"},
           "fused": {"ms_per_frame": ms_mine, "rays_per_s": 65536 / ms_mine * 1e3},
           "speedup": ms_ref / ms_mine,
           "psnr_db": psnr(out[3].reshape(-1, 3).cpu().numpy(), rgb_r.cpu().numpy())}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "reference_gpu.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
