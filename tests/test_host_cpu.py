"""CPU suite for the host-side pieces around the hot path: checkpoint / resume wire format, the asynchronous image
writer, and the refusal of the device-only stages to run without a GPU (no CPU fallback)."""
import os

import numpy as np
import pytest
import torch

from avatarcraft_b200.utils import checkpoint as ckpt


class _Opt:
    def __init__(self):
        self.s = {"step": 7, "lr": 2.5e-3, "betas": (0.9, 0.999), "eps": 1e-8, "exp_avg": torch.arange(8.0), "exp_avg_sq": torch.ones(8)}

    def state_dict(self):
        return self.s

    def load_state_dict(self, sd):
        self.s = sd


def test_checkpoint_is_the_reference_wire_format_plus_resume_file(tmp_path):
    net = torch.nn.Linear(3, 2)
    path = str(tmp_path / "exp" / "exp_0007.pth.tar")
    torch.manual_seed(123)
    rpath = ckpt.save_checkpoint(path, net, _Opt(), step=7, epoch=2, extra={"view": 3})
    expect_next = torch.rand(4)                                   # what the RNG produces right after the save
    # the weights file is exactly torch.save(state_dict) -- what the reference's entry points load (stylize.py:255-260)
    sd = torch.load(path, map_location="cpu")
    assert set(sd.keys()) == {"weight", "bias"} and torch.equal(sd["weight"], net.weight.detach())
    assert rpath == ckpt.resume_path(path) and os.path.exists(rpath)
    net2, opt2 = torch.nn.Linear(3, 2), _Opt()
    opt2.s = None
    torch.manual_seed(999)
    info = ckpt.load_checkpoint(path, net2, opt2)
    assert info == {"step": 7, "epoch": 2, "extra": {"view": 3}, "resumed": True}
    assert torch.equal(net2.weight, net.weight) and opt2.s["step"] == 7 and torch.equal(opt2.s["exp_avg"], torch.arange(8.0))
    assert torch.equal(torch.rand(4), expect_next)                # RNG stream continues where it stopped
    # a plain reference checkpoint (no resume file) loads too
    plain = str(tmp_path / "plain.pth.tar")
    torch.save(net.state_dict(), plain)
    assert ckpt.load_checkpoint(plain, net2, _Opt())["resumed"] is False


def test_async_image_writer_writes_png_and_gif(tmp_path):
    from PIL import Image
    w = ckpt.AsyncImageWriter(gif_path=str(tmp_path / "orbit.gif"))
    frames = [torch.full((8, 12, 3), v) for v in (0.0, 0.5, 1.0)]
    for i, f in enumerate(frames):
        w.submit(f, str(tmp_path / "sub" / f"f_{i}.png"))
    w.close()
    for i, v in enumerate((0, 128, 255)):
        img = np.asarray(Image.open(tmp_path / "sub" / f"f_{i}.png"))
        assert img.shape == (8, 12, 3) and int(img[0, 0, 0]) == v
    assert Image.open(tmp_path / "orbit.gif").n_frames == 3


def test_device_only_stages_refuse_cpu():
    from avatarcraft_b200.utils import ray_gen
    from avatarcraft_b200.utils.optim import FlatAdam
    with pytest.raises(RuntimeError):
        ray_gen.pinhole_rays_device(np.eye(4), 8, 8, "cpu")
    with pytest.raises(RuntimeError):
        FlatAdam([torch.nn.Parameter(torch.zeros(4))])


def test_models_neus_matches_reference_fixture():
    """avatarcraft_b200.models.neus (legacy MLP NeuS API, models/neus.py:647,784) on the CPU against outputs of the reference's
    own module (tests/golden/neus_small.npz, written by oracle/make_golden_neus.py): same state-dict keys, same render."""
    import numpy as np
    import torch
    from tests.util import GOLDEN
    from avatarcraft_b200.models.neus import build_neus, NeuSRenderer
    g = dict(np.load(os.path.join(GOLDEN, "neus_small.npz")))
    neus, params = build_neus(n_sdf=4, n_color=2, w_sdf=48, w_color=32, w_geo_feat=24, skip=[2], use_id=False)
    assert isinstance(neus, NeuSRenderer) and len(params) > 0
    neus = neus.cpu()
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(neus.state_dict())
    neus.load_state_dict(sd)
    o, d, near, far = (torch.from_numpy(g[k]) for k in ("rays_o", "rays_d", "near", "far"))
    for tag, imp in (("coarse", -1), ("fine", 64)):
        r = neus.render(o, d, near, far, perturb_overwrite=0, n_importance_overwrite=imp, background_rgb=torch.ones(1, 3), cos_anneal_ratio=0.7)
        assert set(r) == {"color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients", "weights", "gradient_error",
                          "inside_sphere"}
        for k in ("color_fine", "weight_sum", "weights", "gradient_error", "cdf_fine", "s_val"):
            np.testing.assert_allclose(r[k].detach().numpy(), g[f"{tag}.{k}"], atol=2e-5, rtol=1e-4, err_msg=f"{tag}.{k}")


def test_reference_import_paths_resolve_to_the_package():
    """The reference's entry points import `models.instant_nsr`, `models.neus`, `encoder`, `raymarching`, `utils.render_utils`
    (render_canonical.py:22, stylize.py:26, encoder/__init__.py): with the repo root on sys.path the same statements load this
    package's modules (alias modules at the repo root)."""
    import importlib
    import subprocess
    import sys
    from tests.util import ROOT
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from models.instant_nsr import NeRFNetwork, NeRFRenderer\n"
            "from models.neus import build_neus, NeuSRenderer\n"
            "from models.diffusion import StableDiffusion\n"
            "from encoder import get_encoder\n"
            "from encoder.hashencoder import HashEncoder\n"
            "from encoder.shencoder import SHEncoder\n"
            "import raymarching, utils.render_utils as ru, utils.ray_utils as rays, utils.constant as c\n"
            "import avatarcraft_b200.models.instant_nsr as impl\n"
            "assert NeRFNetwork is impl.NeRFNetwork and sys.modules['models.instant_nsr'] is impl\n"
            "assert hasattr(ru, 'render_instantnsr_naive') and hasattr(rays, 'warp_samples_to_canonical') and hasattr(raymarching, 'march_rays_train')\n"
            "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
