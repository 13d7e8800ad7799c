"""Shared helpers for the parity tests."""
import os
import warnings

import numpy as np
import torch

from avatarcraft_b200.utils import synthetic as syn

warnings.filterwarnings("ignore", category=FutureWarning)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_SD = {}


def state_dict(kind, seed):
    key = (kind, int(seed))
    if key not in _SD:
        _SD[key] = syn.synthetic_state_dict(kind, int(seed))
    return _SD[key]


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    sd = state_dict(str(g["kind"]), int(g["seed"])) if "kind" in g else state_dict("trained", 43)
    assert abs(syn.state_checksum(sd) - float(g["state_checksum"])) < 1e-6 * abs(float(g["state_checksum"])), \
        "synthetic checkpoint differs from the one the golden fixture was generated with (torch RNG drift)"
    return g, sd


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


def gpu_model(sd, train=False, grad=None):
    """Product model on cuda:0.  `grad` (default = `train`) decides whether parameters require grad: with
    autograd on, `run` takes the differentiable path exactly like the reference would build a graph."""
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    net = NeRFNetwork()
    net.load_state_dict(sd)
    net = net.cuda()
    net.train(train)
    for p in net.parameters():
        p.requires_grad_(train if grad is None else grad)
    return net
