"""CPU suite (-m "not gpu"): the oracle against the golden fixtures generated from the
reference's own code, the host-side mirror of the reference API, and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.util import ROOT, load_golden, psnr, state_dict
from oracle import hashgrid as ohg
from oracle.nsr_oracle import OracleNSR
from avatarcraft_b200.utils import synthetic as syn

CASES = ["c1_init_64x64_16p16", "c2_trained_256x256_64p64", "c4_trained_256x256_32p32", "c3_trained_jitter_64p64"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_fixture(name):
    """oracle/nsr_oracle.py == the reference's NeRFRenderer.run (fixture written by
    oracle/make_golden.py from /root/reference) on the same rays and checkpoint."""
    g, sd = load_golden(name)
    jit = torch.from_numpy(g["jitter"]) if g["jitter"].size else None
    out = OracleNSR(sd).run(torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]), int(g["num_steps"]),
                            float(g["bound"]), int(g["upsample_steps"]), jitter=jit)
    depth, weights, wsum, image, nmap, eik, _, color, alpha, z = out
    tol = 2e-6   # same torch build, same op order: agreement is at rounding level
    np.testing.assert_allclose(z.numpy(), g["z_vals"], atol=tol)
    np.testing.assert_allclose(image.reshape(-1, 3).numpy(), g["rgb"], atol=tol)
    np.testing.assert_allclose(weights.numpy(), g["weights"], atol=tol)
    np.testing.assert_allclose(alpha.numpy(), g["pts_alpha"], atol=tol)
    np.testing.assert_allclose(color.numpy(), g["pts_color"], atol=tol)
    np.testing.assert_allclose(nmap.numpy(), g["normal"], atol=tol)
    np.testing.assert_allclose(depth.reshape(-1).numpy(), g["depth"], atol=tol)
    np.testing.assert_allclose(float(eik), float(g["eikonal"]), rtol=1e-5)


def test_hashgrid_oracle_against_reference_wrapper_fixture():
    g, sd = load_golden("hashgrid_trained_768")
    x = torch.from_numpy(g["x"])
    m = OracleNSR(sd)
    feats, ids = m.encode(x, 1.6, want_ids=True)
    np.testing.assert_array_equal(ids.numpy(), g["corner_ids"])
    np.testing.assert_array_equal(feats.numpy(), g["feats"])
    np.testing.assert_allclose(m.forward_sdf(x, 1.6).numpy(), g["sdf16"], atol=1e-6)
    # out-of-range rows (|x| > bound) encode to zero, ids -1 (hashencoder.cu:94-119)
    oob = (x.abs() > 1.6).any(1).numpy()
    assert oob.sum() == 2 and np.all(g["feats"][oob] == 0) and np.all(g["corner_ids"][:, oob] == -1)


def test_hash_offsets_match_reference_table():
    offs, pls = syn.hash_offsets()
    assert offs.tolist() == [0, 4913, 18737, 51505, 136689, 352689, 876977, 1401265, 1925553, 2449841, 2974129,
                             3498417, 4022705, 4546993, 5071281, 5595569, 6119857]
    o2, p2 = ohg.grid_offsets()
    assert o2.tolist() == offs.tolist() and abs(pls - p2) < 1e-15
    sc = ohg.level_scales(16, np.log2(pls), 16)
    assert sc[0] == 15.0 and abs(sc[15] - 2047.0) < 1e-3
    # dense levels are 0..4 (level 4 has 60^3 = 216000 entries), hashed 5..15 (2^19 each)
    sizes = np.diff(offs.numpy())
    assert sizes[4] == 216000 and np.all(sizes[5:] == 2 ** 19)


def test_oracle_backward_matches_autograd_of_forward():
    """The table gradient of the C oracle equals d(sum(out*g))/d(table) by finite linearity:
    the encoder is linear in the table, so backward(g) . t == forward(t) . g for any t."""
    torch.manual_seed(3)
    offs, pls = ohg.grid_offsets(num_levels=4, log2_hashmap_size=10, desired_resolution=64)
    offs_t = torch.from_numpy(offs)
    n = int(offs[-1])
    B, L, C, D = 300, 4, 2, 3
    x = torch.rand(B, D)
    g = torch.randn(L, B, C)
    t = torch.randn(n, C)
    S = float(np.log2(pls))
    out = torch.empty(L, B, C)
    ohg.hash_encode_forward(x, t, offs_t, out, B, D, C, L, S, 16, False, None)
    gt = torch.zeros(n, C)
    ohg.hash_encode_backward(g, x, t, offs_t, gt, B, D, C, L, S, 16, False, None, None)
    lhs, rhs = float((gt.double() * t.double()).sum()), float((out.double() * g.double()).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(rhs))


def test_state_dict_layout_matches_reference():
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    net = NeRFNetwork()
    sd = net.state_dict()
    assert tuple(sd.keys()) == syn.STATE_KEYS
    ref = state_dict("init", 42)
    for k in syn.STATE_KEYS:
        assert tuple(sd[k].shape) == tuple(ref[k].shape) and sd[k].dtype == ref[k].dtype, k
    assert sum(v.numel() for v in sd.values()) == 12248919
    net.load_state_dict(ref)                                           # reference-layout checkpoints load as is


def test_c_abi_exports_every_declared_symbol():
    """libavatarcraft_b200.so loads without a GPU and exports what include/*.h declares."""
    from avatarcraft_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "avatarcraft_b200.h")).read()
    declared = set(re.findall(r"\b(ac_[a-z0-9_]+)\s*\(", hdr))
    handle = ctypes.CDLL(_lib.build())
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) <= declared
    assert b"sm_100a" in _lib.lib().ac_version()


def test_product_path_has_no_cpu_fallback():
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    net = NeRFNetwork()
    o, d = syn.pinhole_rays(syn.orbit_pose(0.0), 4, 4)
    with pytest.raises(RuntimeError):
        net.render(o[None], d[None], num_steps=16, bound=1.6, upsample_steps=16)
    with pytest.raises(RuntimeError):
        net.encoder(torch.zeros(4, 3), 1.6)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "avatarcraft_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dp, f)


def test_synthetic_rays_match_reference_camera():
    """Pinned against the value obtained by running the reference's cameras/ + shot_rays code
    (SURVEY.md 8c): orbit view at +z, 8x8 image."""
    o, d = syn.pinhole_rays(syn.orbit_pose(0.0), 8, 8)
    np.testing.assert_allclose(o[0].numpy(), [0, 0, 1.7], atol=1e-6)
    np.testing.assert_allclose(d[0].numpy(), [-0.4745, 0.4745, -0.7414], atol=1e-4)
    assert o.dtype == torch.float32 and abs(float(d.norm(dim=1).mean()) - 1) < 1e-6


def training_loss(out, pixel_grad):
    """Same three terms as oracle/make_golden.py::training_loss (stylize.py:163-193)."""
    rgb, wsum, eik = out[3].reshape(-1, 3), out[2], out[5]
    opacity = torch.nn.functional.smooth_l1_loss(wsum.clamp(0.0, 1.0), torch.full_like(wsum, 0.5)) * 1e5
    return (rgb * pixel_grad).sum() + 0.01 * eik + opacity * 1e-3


def test_oracle_gradients_match_reference_autograd_fixture():
    """Parameter gradients of the differentiable oracle == the reference's own autograd
    (fixture written from /root/reference by oracle/make_golden.py), training mode with jitter."""
    g, sd = load_golden("grad_trained_jitter_64p64")
    orc = OracleNSR(sd)
    params = orc.enable_grad(sd)
    out = orc.run_grad(torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]), 64, 1.6, 64,
                       jitter=torch.from_numpy(g["jitter"]))
    loss = training_loss(out, torch.from_numpy(g["pixel_grad"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    for k, p in params.items():
        if k == "encoder.embeddings":
            rows = torch.from_numpy(g["emb_rows"])
            np.testing.assert_allclose(p.grad[rows].numpy(), g["emb_grad"], atol=1e-8, rtol=1e-4)
            assert abs(float(p.grad.abs().double().sum()) - float(g["emb_grad_abs_sum"])) < 1e-5 * float(g["emb_grad_abs_sum"])
            assert int((p.grad.abs().sum(1) > 0).sum()) == int(g["emb_grad_nnz"])
        else:
            ref = g["g." + k]
            np.testing.assert_allclose(p.grad.numpy(), ref, atol=2e-5 * max(1.0, float(np.abs(ref).max())), rtol=1e-4)


def test_warp_oracle_geometry():
    """Closed-form checks of the restated closest-point / warp (igl is unavailable, so these pin the
    restatement to geometry, not to igl): points on the surface have distance 0; a point above a face
    centroid projects onto it with barycentrics (1/3,1/3,1/3); identity-like transforms scale by 0.9."""
    from oracle import warp_oracle as wo
    body = syn.synthetic_body()
    assert body["world_verts"].shape == (6890, 3) and body["faces"].shape == (13776, 6) and body["Ts"].shape == (6914, 4, 4)
    V, F = body["rest_verts"].astype(np.float64), body["faces"][:, :3]
    tri = V[F[100:140]]
    cen = tri.mean(1)
    nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d2, face, closest, bary = wo.closest_point_on_mesh(cen + 0.01 * nrm, V, F)
    np.testing.assert_allclose(d2, 1e-4, rtol=1e-6)
    np.testing.assert_allclose(closest, cen, atol=1e-12)
    assert (face == np.arange(100, 140)).all() and np.allclose(bary, 1 / 3)
    d2v, *_ = wo.closest_point_on_mesh(V[:64], V, F)
    assert d2v.max() < 1e-24
    T = np.tile(np.eye(4) / 0.9, (6914, 1, 1))
    can, mask, *_ = wo.warp_samples_to_canonical((cen + 0.01 * nrm)[None], V, body["faces"], T, 0.05)
    np.testing.assert_allclose(can[0], 0.9 * (cen + 0.01 * nrm), atol=1e-12)       # inverse of diag(1/0.9): xyz * 0.9, w ignored
    assert mask.all()


def test_masked_samples_cannot_reach_the_warped_image():
    """The premise of NeRFNetwork.warp_skip_masked, checked on the oracle (models/instant_nsr.py:245-248): alpha is multiplied by
    the warp's mask, so replacing the canonical position of every masked-out SECTION sample by anything finite (here: the
    un-warped point, which is what ac_warp_samples_to_canonical_masked returns) leaves image, depth, opacity, normals and
    weights bit-identical."""
    from oracle import warp_oracle as wo
    body = syn.synthetic_body()
    orc = OracleNSR(state_dict("trained", 43))
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 64, 64)
    sel = torch.cat([torch.arange(64 * 30 + 20, 64 * 30 + 44, 3), torch.arange(64 * 30, 64 * 30 + 6)])     # torso rays + empty rays
    o, d = o[sel].contiguous(), d[sel].contiguous()
    kw = dict(verts=body["world_verts"], faces=body["faces"], Ts=body["Ts"])
    exact = orc.run(o, d, 32, 1.6, 32, **kw)
    real, calls, masked = wo.warp_samples_to_canonical, [0], [0]

    def bounded(p, *a, **k):
        out = list(real(p, *a, **k))
        calls[0] += 1
        if calls[0] == 2:                           # the section-point warp (the coarse warp feeds the importance sampling: exact)
            off = ~np.asarray(out[1], bool)
            masked[0] = int(off.sum())
            out[0] = np.where(off[..., None], np.asarray(p, np.float64), out[0])
        return tuple(out)
    wo.warp_samples_to_canonical = bounded
    try:
        fast = orc.run(o, d, 32, 1.6, 32, **kw)
    finally:
        wo.warp_samples_to_canonical = real
    assert calls[0] == 2 and masked[0] > 100
    wsum = exact[2].reshape(-1)
    assert float(wsum.max()) > 0.5 and float(wsum.min()) == 0.0                 # body rays and fully masked rays both present
    for i, name in ((0, "depth"), (1, "weights"), (2, "weights_sum"), (3, "image"), (4, "normal_map"), (8, "alpha"), (9, "z_vals")):
        assert torch.equal(exact[i], fast[i]), name


def test_smpl_lbs_matches_reference_fixture():
    """avatarcraft_b200.models.smpl (host-side W3) == the reference's models/smpl.py::lbs on the synthetic
    SMPL-shaped model (fixture from /root/reference): per-vertex + joint transforms, posed vertices, joints."""
    from avatarcraft_b200.models.smpl import SMPL, calc_local_trans
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "smpl_lbs_synthetic.npz")))
    m = SMPL(syn.synthetic_smpl_model())
    _, T, _ = m.verts_transformations(g["pose"], g["betas"], concat_joints=True)
    np.testing.assert_allclose(T[0, g["T_rows"]].numpy(), g["T"], atol=1e-6)
    v, J = m.forward(g["pose"], g["betas"], return_joints=True)
    np.testing.assert_allclose(v[0, g["posed_rows"]].numpy(), g["posed"], atol=1e-6)
    np.testing.assert_allclose(J[0].numpy(), g["joints"], atol=1e-6)
    wv, Ts, n = calc_local_trans(m, poses=syn.sinusoid_pose_sequence(3))
    assert n == 3 and wv[0].shape == (6890, 3) and Ts[0].shape == (6914, 4, 4)
    assert np.abs(Ts[1][:, 3, :3]).max() == 0 and np.allclose(Ts[1][:, 3, 3], 1 / 0.9)      # (0,0,0,1/0.9) rows (render_warp.py:200-204)
    # the posed surface is the rest surface pushed through T_rest2pose: inverse warp of a posed vertex lands on the rest vertex * 0.9
    rest, _ = m.forward(np.zeros((1, 72), np.float32) + _da_pose(), np.zeros((1, 10), np.float32), return_joints=True)
    back = np.einsum("nij,nj->ni", np.linalg.inv(Ts[1][:6890]), np.concatenate([wv[1], np.ones((6890, 1))], 1))[:, :3]
    np.testing.assert_allclose(back, 0.9 * rest[0].numpy(), atol=2e-5)


def _da_pose():
    da = np.zeros((24, 3), np.float32)
    da[1], da[2] = [0, 0, 1.0], [0, 0, -1.0]
    return da.reshape(1, 72)


def test_gen_rays_pose_convention():
    from avatarcraft_b200.utils.ray_gen import dataset_intrinsics, gen_rays_pose
    K = dataset_intrinsics()
    pose = np.eye(4, dtype=np.float32); pose[:3, 3] = [0.1, 0.2, 2.0]
    o, v = gen_rays_pose(pose, K, 512, 512, resolution_level=2)
    assert o.shape == (256, 256, 3) and v.shape == (256, 256, 3)
    np.testing.assert_allclose(o[5, 7].numpy(), [0.1, 0.2, 2.0])
    np.testing.assert_allclose(torch.linalg.norm(v, dim=-1).numpy(), 1.0, atol=1e-6)
    assert float(v[128, 128, 2]) < -0.99 and float(v[0, 0, 0]) < 0 < float(v[0, 0, 1])      # looks down -z, +y is up, x grows right
