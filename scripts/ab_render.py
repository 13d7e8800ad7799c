"""A/B of the two fused render kernels on the GPU box: the stencil-sharing kernel (default) against the round-1 kernel
(AC_RENDER_IMPL=tc5) on the BASELINE frame (256x256, 64+64): output differences and ms/frame (L2 flushed, CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatarcraft_b200.models.instant_nsr import NeRFNetwork  # noqa: E402
from avatarcraft_b200.utils import synthetic as syn  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    net = NeRFNetwork(); net.load_state_dict(syn.synthetic_state_dict("trained", 43)); net = net.to(dev).eval()
    o, d = syn.pinhole_rays(syn.orbit_pose(30.0), 256, 256)
    o, d = o.to(dev), d.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.set_grad_enabled(False)
    res = {}

    def run(impl, per_sample):
        if impl:
            os.environ["AC_RENDER_IMPL"] = impl
        else:
            os.environ.pop("AC_RENDER_IMPL", None)
        return net.run(o[None], d[None], 64, 1.6, 64, None, 1.0, 0.0, per_sample_outputs=per_sample)

    def timed(impl, n=20):
        for _ in range(3):
            run(impl, False)
        tot = 0.0
        best = 1e9
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(impl, False); e1.record(); e1.synchronize()
            t = e0.elapsed_time(e1); tot += t; best = min(best, t)
        return tot / n, best

    new = run("st", True); torch.cuda.synchronize()
    old = run("tc5", True); torch.cuda.synchronize()
    names = ["depth", "weights", "weight_sum", "rgb", "normal", "eikonal", None, "pts_color", "pts_alpha", "z_vals"]
    for nm, a, b in zip(names, new, old):
        if nm is None:
            continue
        a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
        res[nm] = {"max_abs": float((a - b).abs().max()), "equal_frac": float((a == b).float().mean())}
    mse = float(((new[3] - old[3]) ** 2).mean())
    res["psnr_new_vs_old"] = 99.0 if mse == 0 else float(10 * torch.log10(torch.tensor(1.0 / mse)))
    res["ms_new"], res["ms_new_best"] = timed("st")
    res["ms_old"], res["ms_old_best"] = timed("tc5")
    os.environ.pop("AC_RENDER_IMPL", None)
    # small launches (a 512-ray shard of a training patch) and the sampling-only launch
    for n in (512, 4096):
        oo, dd = o[32768:32768 + n].contiguous(), d[32768:32768 + n].contiguous()
        for impl in ("st", "tc5"):
            if impl:
                os.environ["AC_RENDER_IMPL"] = impl
            else:
                os.environ.pop("AC_RENDER_IMPL", None)
            for _ in range(3):
                net.run(oo[None], dd[None], 64, 1.6, 64, None, 1.0, 0.0, per_sample_outputs=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                net.run(oo[None], dd[None], 64, 1.6, 64, None, 1.0, 0.0, per_sample_outputs=False)
            e1.record(); e1.synchronize()
            res[f"ms_{n}rays_{impl or 'st'}"] = e0.elapsed_time(e1) / 10
    os.environ.pop("AC_RENDER_IMPL", None)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
