"""oracle/make_golden_neus.py -- TEST INFRASTRUCTURE ONLY; runs in the BUILD container only.

Pins avatarcraft_b200/models/neus.py (the legacy MLP NeuS API) to the reference: imports the reference's own
models/neus.py from /root/reference, builds a small model with build_neus under a fixed seed, renders a seeded ray batch on
the CPU (with and without importance sampling) and writes the state-dict + inputs + outputs to tests/golden/neus_small.npz.

    python -m oracle.make_golden_neus
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, GOLD     # noqa: E402

CFG = dict(n_sdf=4, n_color=2, w_sdf=48, w_color=32, w_geo_feat=24, skip=[2], use_id=False)


def main():
    import_reference()
    import models.neus as ref_neus
    torch.manual_seed(7)
    neus, _ = ref_neus.build_neus(**CFG)
    neus = neus.cpu()
    gen = torch.Generator().manual_seed(8)
    n = 96
    o = torch.randn(n, 3, generator=gen) * 0.1 + torch.tensor([0.0, 0.0, 2.0])
    d = -o + torch.randn(n, 3, generator=gen) * 0.3
    d = d / d.norm(dim=-1, keepdim=True)
    near, far = torch.full((n, 1), 1.0), torch.full((n, 1), 3.0) + torch.rand(n, 1, generator=gen) * 0.2      # per-ray bounds [N,1]
    out = {}
    for tag, imp in (("coarse", -1), ("fine", 64)):
        r = neus.render(o, d, near, far, perturb_overwrite=0, n_importance_overwrite=imp, background_rgb=torch.ones(1, 3),
                        cos_anneal_ratio=0.7)
        for k in ("color_fine", "weight_sum", "weights", "gradient_error", "cdf_fine", "s_val"):
            out[f"{tag}.{k}"] = r[k].detach().numpy()
    sd = {f"sd.{k}": v.detach().numpy() for k, v in neus.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, "neus_small.npz"), rays_o=o.numpy(), rays_d=d.numpy(), near=near.numpy(), far=far.numpy(), **sd, **out)
    print("wrote neus_small.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
