"""oracle/make_golden_viewdirs.py -- TEST INFRASTRUCTURE ONLY; BUILD container only.

use_viewdirs=True (models/instant_nsr.py:564-569,646-650) is dormant in the reference's entry points but part of NeRFNetwork's
API.  This script runs the REFERENCE'S OWN NeRFNetwork(use_viewdirs=True).run on the CPU -- hash backend = oracle/hashgrid.py,
SH backend = scipy's spherical harmonics (oracle/sh_oracle.py; the ray directions are unit vectors) -- and writes
tests/golden/c6_viewdirs_1280rays_32p32.npz.      python -m oracle.make_golden_viewdirs"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, GOLD, pack     # noqa: E402
from oracle import sh_oracle                                     # noqa: E402
from avatarcraft_b200.utils import synthetic as syn              # noqa: E402


def main():
    ref = import_reference()

    def sh_forward(inputs, outputs, B, D, degree, calc_grad, dy_dx):
        outputs.copy_(torch.from_numpy(sh_oracle.sh_scipy(inputs.detach().double().numpy(), degree)).float())
    sys.modules["encoder.shencoder.backend"]._backend = types.SimpleNamespace(sh_encode_forward=sh_forward)
    import encoder.shencoder.sphere_harmonics as shmod
    shmod._backend = sys.modules["encoder.shencoder.backend"]._backend
    sd = syn.synthetic_state_dict("trained", 43)
    net = ref.NeRFNetwork(use_viewdirs=True)
    gen = torch.Generator().manual_seed(61)
    c0_v = torch.randn(64, 37, generator=gen) * 0.35
    c0_g = c0_v.norm(dim=1, keepdim=True) * (0.8 + 0.4 * torch.rand(64, 1, generator=gen))
    sd = dict(sd)
    sd["color_net.0.weight_v"], sd["color_net.0.weight_g"] = c0_v, c0_g
    net.load_state_dict(sd)
    net.eval()
    o, d = syn.pinhole_rays(syn.orbit_pose(40.0), 64, 64)
    sel = torch.arange(64 * 12, 64 * 52, 2)                     # 1280 rays through the body
    o, d = o[sel].contiguous(), d[sel].contiguous()
    with torch.no_grad():
        out = net.run(o[None], d[None], 32, 1.6, 32, None, cos_anneal_ratio=1.0, normal_epsilon_ratio=0.0, render_can=True)
    np.savez_compressed(os.path.join(GOLD, "c6_viewdirs_1280rays_32p32.npz"), rays_o=o.numpy(), rays_d=d.numpy(), num_steps=32, upsample_steps=32,
                        c0_v=c0_v.numpy(), c0_g=c0_g.numpy(), kind="trained", seed=43, state_checksum=syn.state_checksum(syn.synthetic_state_dict("trained", 43)),
                        **{k: v for k, v in pack(out).items() if k in ('rgb', 'depth', 'weight_sum', 'normal', 'eikonal', 'z_vals', 'pts_color')})
    print("wrote c6_viewdirs_1280rays_32p32.npz; hit fraction", float((out[2] > 0.5).float().mean()))


if __name__ == "__main__":
    main()
