"""GPU suite for the spherical-harmonics encoder op (dormant in the reference unless
use_viewdirs=True, models/instant_nsr.py:564-569).  Oracle: oracle/sh_oracle.py (closed forms from the
reference's own comments for degree <= 3, scipy's complex harmonics on unit vectors for all 64)."""
import numpy as np
import pytest
import torch

from oracle import sh_oracle as so

pytestmark = pytest.mark.gpu


def _encode(x, degree, jac=True):
    from avatarcraft_b200.encoder.shencoder.backend import _backend
    B = x.shape[0]
    out = torch.empty(B, degree * degree, device="cuda")
    dy_dx = torch.empty(B, 3 * degree * degree, device="cuda")
    _backend.sh_encode_forward(x, out, B, 3, degree, jac, dy_dx)
    torch.cuda.synchronize()
    return out, dy_dx.reshape(B, 3, degree * degree)


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_forward_against_scipy_on_unit_vectors(degree):
    p = np.random.default_rng(degree).normal(size=(4000, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    p[:3] = np.eye(3)                                                     # poles / axes
    out, _ = _encode(torch.from_numpy(p).float().cuda(), degree)
    ref = so.sh_scipy(p, degree)
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=3e-5, rtol=2e-5)


def test_sh_non_unit_inputs_and_jacobian_against_closed_forms():
    """Arbitrary (non-normalised) xyz: the polynomials of shencoder.cu:48-58; Jacobian vs float64 central differences."""
    p = np.random.default_rng(1).uniform(-1.5, 1.5, size=(3000, 3))
    out, jac = _encode(torch.from_numpy(p).float().cuda(), 3)
    np.testing.assert_allclose(out.cpu().numpy(), so.sh_closed_form(p), atol=5e-6, rtol=1e-5)
    h = 1e-6
    for d in range(3):
        e = np.zeros(3); e[d] = h
        fd = (so.sh_closed_form(p + e) - so.sh_closed_form(p - e)) / (2 * h)
        np.testing.assert_allclose(jac[:, d].cpu().numpy(), fd, atol=2e-5, rtol=1e-5)


def test_sh_jacobian_consistent_with_forward_degree8_and_backward():
    """Degree 8: Jacobian vs central differences of the op's own forward (fp32, h=2e-3), and the backward op
    grad_inputs[b,d] = sum_c grad[b,c] dy_dx[b,d,c] (shencoder.cu:360-384)."""
    from avatarcraft_b200.encoder.shencoder.backend import _backend
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(2000, 3, generator=g) * 2 - 1).cuda()
    out, jac = _encode(x, 8)
    h = 2e-3
    for d in range(3):
        e = torch.zeros(3, device="cuda"); e[d] = h
        fd = (_encode(x + e, 8, False)[0] - _encode(x - e, 8, False)[0]) / (2 * h)
        scale = float(jac[:, d].abs().max())
        assert float((jac[:, d] - fd).abs().max()) < 5e-3 * scale
    grad = torch.randn(2000, 64, generator=g).cuda()
    gi = torch.zeros(2000, 3, device="cuda")
    _backend.sh_encode_backward(grad, x, 2000, 3, 8, jac.reshape(2000, -1).contiguous(), gi)
    np.testing.assert_allclose(gi.cpu().numpy(), torch.einsum("bc,bdc->bd", grad, jac).cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_sh_module_and_factory():
    from avatarcraft_b200.encoder import get_encoder
    enc, dim = get_encoder("sphere_harmonics", {"in_dim": 3})
    assert dim == 16
    d = torch.nn.functional.normalize(torch.randn(100, 3), dim=-1).cuda().requires_grad_(True)
    y = enc(d)
    y.sum().backward()
    assert y.shape == (100, 16) and d.grad.shape == (100, 3) and torch.isfinite(d.grad).all()
    np.testing.assert_allclose(y.detach().cpu().numpy(), so.sh_scipy(d.detach().cpu().numpy().astype(np.float64), 4), atol=2e-5)
    with pytest.raises(RuntimeError):
        from avatarcraft_b200.encoder.shencoder.backend import _backend
        _backend.sh_encode_forward(d.detach(), y.detach(), 100, 3, 9, False, y.detach())     # degree 9 unsupported


def test_point_major_hash_ops_equal_reference_layout_ops():
    """ac_hash_encode_forward_pm / _backward_pm ([B, L*C] in and out) == the reference-layout ops up to the permute;
    HashEncoder (module) gradients flow to the table and, when asked, to the inputs."""
    from avatarcraft_b200 import _lib
    from avatarcraft_b200.encoder.hashencoder import GridSpec, HashEncoder, hash_encode
    from avatarcraft_b200.encoder.hashencoder.backend import _backend
    spec = GridSpec(3, 16, 2, 16, float(np.exp2(np.log2(2048 / 16) / 15)), 19)
    offs = torch.from_numpy(spec.level_offsets()).cuda()
    g = torch.Generator().manual_seed(2)
    table = ((torch.rand(int(offs[-1]), 2, generator=g) * 2 - 1) * 0.1).cuda().requires_grad_(True)
    x = torch.rand(20000, 3, generator=g).cuda().requires_grad_(True)
    feats = hash_encode(x, table, offs, spec)
    ref = torch.empty(16, 20000, 2, device="cuda"); jac = torch.empty(20000, 16 * 3 * 2, device="cuda")
    _backend.hash_encode_forward(x.detach(), table.detach(), offs, ref, 20000, 3, 2, 16, spec.log2_scale, 16, True, jac)
    assert torch.equal(feats.detach(), ref.permute(1, 0, 2).reshape(20000, 32))
    R = torch.randn(20000, 32, generator=g).cuda()
    (feats * R).sum().backward()
    gt = torch.zeros_like(table); gi = torch.zeros(20000, 3, device="cuda")
    _backend.hash_encode_backward(R.view(20000, 16, 2).permute(1, 0, 2).contiguous(), x.detach(), table.detach(), offs, gt, 20000, 3, 2, 16,
                                  spec.log2_scale, 16, True, jac, gi)
    np.testing.assert_allclose(table.grad.cpu().numpy(), gt.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), gi.cpu().numpy(), rtol=1e-4, atol=1e-2)
    enc = HashEncoder(desired_resolution=2048, per_level_scale=1.3819).cuda()
    assert enc.offsets.tolist() == spec.level_offsets().tolist() and enc.output_dim == 32
    y = enc((torch.rand(100, 7, 3, device="cuda") * 2 - 1) * 1.6, 1.6)
    y.square().sum().backward()
    assert y.shape == (100, 7, 32) and enc.embeddings.grad is not None and float(enc.embeddings.grad.abs().sum()) > 0


def _ref_sh_module():
    """The reference's own shencoder.cu compiled for sm_100a by oracle/build_ref.py (test infrastructure)."""
    import importlib.machinery, importlib.util, os
    from tests.util import ROOT
    so_path = os.path.join(ROOT, "oracle", "_ref", "_ref_sh_encoder.so")
    if not os.path.exists(so_path):
        pytest.skip("oracle/_ref/_ref_sh_encoder.so not built (run oracle/build_ref.py in the build container)")
    loader = importlib.machinery.ExtensionFileLoader("_ref_sh_encoder", so_path)
    spec = importlib.util.spec_from_loader("_ref_sh_encoder", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_ops_against_reference_kernel(degree):
    """ac_sh_encode_forward / _backward next to the REFERENCE'S OWN kernel_sh / kernel_sh_backward (shencoder.cu:28-384) on the
    same inputs (unit and non-unit directions).  The reference evaluates 64 hard-coded polynomials, this library a Legendre
    recurrence: same polynomials, different association, so values agree to fp32 rounding (1e-6 relative to the largest
    coefficient of the degree), not bit for bit."""
    from avatarcraft_b200.encoder.shencoder.backend import _backend
    ref = _ref_sh_module()
    B, C = 50000, degree * degree
    g = torch.Generator().manual_seed(100 + degree)
    x = torch.randn(B, 3, generator=g)
    x[: B // 2] /= x[: B // 2].norm(dim=-1, keepdim=True)                  # half unit directions, half arbitrary points in ~[-3, 3]
    x[:3] = torch.eye(3)
    x = x.cuda().contiguous()
    out_r = torch.empty(B, C, device="cuda"); jac_r = torch.empty(B, 3 * C, device="cuda")
    out_m = torch.empty_like(out_r); jac_m = torch.empty_like(jac_r)
    ref.sh_encode_forward(x, out_r, B, 3, degree, True, jac_r)
    _backend.sh_encode_forward(x, out_m, B, 3, degree, True, jac_m)
    torch.cuda.synchronize()
    so_, sj = float(out_r.abs().max()), float(jac_r.abs().max())
    eo, ej = float((out_m - out_r).abs().max()), float((jac_m - jac_r).abs().max())
    print(f"SH degree {degree}: max |out - ref| = {eo:.2e} (scale {so_:.2e}), max |dy_dx - ref| = {ej:.2e} (scale {sj:.2e})")
    # unit half: coefficients are O(1), absolute agreement to a few ulps
    eu = float((out_m[: B // 2] - out_r[: B // 2]).abs().max())
    print(f"             unit directions only: max |out - ref| = {eu:.2e}")
    assert eu <= 1e-6 * degree * degree                                      # O(1) coefficients: a few ulps per recurrence step
    assert eo <= 1e-6 * max(so_, 1.0) and ej <= 1e-6 * max(sj, 1.0)
    grad = torch.randn(B, C, generator=g).cuda()
    gi_r = torch.zeros(B, 3, device="cuda"); gi_m = torch.zeros(B, 3, device="cuda")
    ref.sh_encode_backward(grad, x, B, 3, degree, jac_r, gi_r)
    _backend.sh_encode_backward(grad, x, B, 3, degree, jac_r, gi_m)          # same Jacobian in: the contraction alone
    torch.cuda.synchronize()
    np.testing.assert_allclose(gi_m.cpu().numpy(), gi_r.cpu().numpy(), atol=1e-5 * max(float(gi_r.abs().max()), 1.0), rtol=1e-5)
