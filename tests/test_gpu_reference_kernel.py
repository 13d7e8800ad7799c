"""GPU pin of the hash encoder against the REFERENCE'S OWN CUDA kernel.

oracle/_ref/_ref_hash_encoder.so is the reference's unmodified encoder/hashencoder/src/
{hashencoder.cu,bindings.cpp} compiled for sm_100a by oracle/build_ref.py (only -std=c++14 ->
c++17).  It is test infrastructure: the checker, never the thing shipped.  Integer path and
forward values must be BIT-IDENTICAL -- same expressions, same FMA contractions, same exp2f."""
import importlib.machinery
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.util import ROOT, state_dict
from oracle import hashgrid as ohg

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_ref_hash_encoder.so")


def ref_module():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built (run oracle/build_ref.py in the build container)")
    loader = importlib.machinery.ExtensionFileLoader("_ref_hash_encoder", REF_SO)
    spec = importlib.util.spec_from_loader("_ref_hash_encoder", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("D,C,L,log2T,res,B", [(3, 2, 16, 19, 2048, 200000), (2, 2, 8, 12, 256, 30000),
                                               (3, 4, 6, 14, 128, 30000), (3, 1, 6, 14, 128, 30000)])
def test_forward_bit_identical_to_reference_kernel(D, C, L, log2T, res, B):
    from avatarcraft_b200.encoder.hashencoder.backend import _backend
    ref = ref_module()
    torch.manual_seed(D * 10 + C)
    offs, pls = ohg.grid_offsets(D, L, None, 16, log2T, res)
    S = float(np.log2(pls))
    n = int(offs[-1])
    x = torch.rand(B, D, device="cuda")
    x[:3] = torch.tensor([[0.0] * D, [1.0] * D, [0.5] * D], device="cuda")
    x[3] = 1.5                                            # out of range
    table = (torch.rand(n, C, device="cuda") * 2 - 1)
    od = torch.from_numpy(offs).cuda()
    out_r = torch.empty(L, B, C, device="cuda"); jac_r = torch.empty(B, L * D * C, device="cuda")
    out_m = torch.empty_like(out_r); jac_m = torch.empty_like(jac_r)
    ref.hash_encode_forward(x, table, od, out_r, B, D, C, L, S, 16, True, jac_r)
    _backend.hash_encode_forward(x, table, od, out_m, B, D, C, L, S, 16, True, jac_m)
    torch.cuda.synchronize()
    assert torch.equal(out_m, out_r), float((out_m - out_r).abs().max())
    assert torch.equal(jac_m, jac_r), float((jac_m - jac_r).abs().max())
    # backward: both scatter with fp32 atomics in unspecified order
    g = torch.randn(L, B, C, device="cuda")
    gt_r = torch.zeros(n, C, device="cuda"); gi_r = torch.zeros(B, D, device="cuda")
    gt_m = torch.zeros(n, C, device="cuda"); gi_m = torch.zeros(B, D, device="cuda")
    ref.hash_encode_backward(g, x, table, od, gt_r, B, D, C, L, S, 16, True, jac_r, gi_r)
    _backend.hash_encode_backward(g, x, table, od, gt_m, B, D, C, L, S, 16, True, jac_m, gi_m)
    torch.cuda.synchronize()
    scale = float(gt_r.abs().max())
    assert float((gt_m - gt_r).abs().max()) <= 1e-5 * scale + 1e-5
    np.testing.assert_allclose(gi_m.cpu().numpy(), gi_r.cpu().numpy(), rtol=1e-4, atol=1e-2)


def test_model_checkpoint_features_bit_identical():
    """The Instant-NSR table (L=16, C=2, 6.1 M entries) on points drawn inside the scene cube."""
    from avatarcraft_b200.encoder.hashencoder.backend import _backend
    ref = ref_module()
    sd = state_dict("trained", 43)
    table, od = sd["encoder.embeddings"].cuda(), sd["encoder.offsets"].cuda()
    B = 524288                                            # one FD pass of a 4096-ray batch
    x = torch.rand(B, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    S = float(np.log2(np.exp2(np.log2(2048 / 16) / 15)))
    out_r = torch.empty(16, B, 2, device="cuda"); out_m = torch.empty_like(out_r)
    dummy = torch.empty(1, device="cuda")
    ref.hash_encode_forward(x, table, od, out_r, B, 3, 2, 16, S, 16, False, dummy)
    _backend.hash_encode_forward(x, table, od, out_m, B, 3, 2, 16, S, 16, False, dummy)
    assert torch.equal(out_m, out_r)
