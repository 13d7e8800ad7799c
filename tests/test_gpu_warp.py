"""GPU suite for the SMPL-guided inverse warp path (BASELINE config 4 shapes).

The closest-point query of the reference is libigl (not installed, not in /root/reference): the oracle
restates it from the geometric definition, so parity with igl's tie-breaking is UNPINNED; distances,
barycentric blends, inverses and everything downstream are compared against the float64 oracle.
Tolerances: fp32 kernel vs float64 numpy -> 2e-4 on canonical points (|T^-1| ~ 1, coordinates ~ 1),
exact mask except within 1e-6 of the threshold."""
import numpy as np
import pytest
import torch

from tests.util import psnr, state_dict, gpu_model
from oracle import warp_oracle as wo
from oracle.nsr_oracle import OracleNSR
from avatarcraft_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def body():
    return syn.synthetic_body()


def test_warp_samples_to_canonical_against_float64_oracle(body):
    from avatarcraft_b200.utils.ray_utils import warp_samples_to_canonical
    gen = torch.Generator().manual_seed(3)
    v = torch.from_numpy(body["world_verts"])
    near = v[torch.randint(0, 6890, (1500,), generator=gen)] + torch.randn(1500, 3, generator=gen) * 0.08
    far = (torch.rand(500, 3, generator=gen) * 2 - 1) * 1.2
    on = v[:48].clone()                                                  # exactly on vertices: dist 0, vertex regions
    pts = torch.cat([near, far, on]).reshape(64, 32, 3)
    can_o, mask_o, closest_o, face_o, d2_o = wo.warp_samples_to_canonical(pts.numpy(), body["world_verts"], body["faces"], body["Ts"], 0.05)
    can, dirs, closest, mask, face, d2 = warp_samples_to_canonical(pts.cuda(), body["world_verts"], body["faces"], body["Ts"], 0.05,
                                                                  return_query=True)
    np.testing.assert_allclose(d2.cpu().numpy(), d2_o, atol=2e-6, rtol=1e-4)
    face_d, same_face = face.cpu().numpy(), face.cpu().numpy() == face_o
    # Off a closed surface the closest point is on an edge or a vertex about a third of the time (their Voronoi regions have
    # volume), i.e. it belongs to two or more faces at the same distance: the face INDEX is then a tie (igl's own choice is
    # unknown anyway) while the closest point -- and with it the barycentric blend of the per-vertex transforms, which only
    # involves the shared vertices -- is unique.  So: the point itself must agree everywhere, and every differing index must
    # be a face that shares a vertex with the oracle's.
    print(f"closest-point query: face index identical on {same_face.mean():.4f} of {same_face.size} points")
    F = np.asarray(body["faces"])
    shares = np.array([len(set(F[a]) & set(F[b])) > 0 for a, b in zip(face_d[~same_face].ravel(), face_o[~same_face].ravel())])
    print(f"  differing indices that are incident faces (share a vertex with the oracle's face): {shares.mean():.4f}")
    assert shares.all()
    dc = np.abs(closest.cpu().numpy() - closest_o).max(-1)
    print(f"  closest point within 5e-5 of the oracle's on {(dc <= 5e-5).mean():.4f} of all points (max {dc.max():.2e})")
    assert (dc <= 5e-5).mean() >= 0.999          # a handful may sit exactly between two different spots of a concave fold
    np.testing.assert_allclose(can.cpu().numpy()[dc <= 5e-5], can_o[dc <= 5e-5], atol=2e-4)
    assert np.abs(can.cpu().numpy() - can_o).max() < 5e-3                 # blended transform is continuous across ties
    safe = np.abs(d2_o - 0.05) > 1e-6
    assert (mask.cpu().numpy() == mask_o)[safe].all()
    assert dirs.shape == (64, 32, 3) and torch.isfinite(dirs[mask]).all()


def test_morton_ordered_queries_equal_plain_order(body):
    """The search is exact, so visiting the queries in Morton order (large batches) must not change a single result."""
    from avatarcraft_b200.utils import ray_utils as ru
    gen = torch.Generator().manual_seed(12)
    v = torch.from_numpy(np.asarray(body["world_verts"], np.float32))
    pts = (v[torch.randint(0, 6890, (40000,), generator=gen)] + torch.randn(40000, 3, generator=gen) * 0.15).reshape(1250, 32, 3).cuda()
    mesh = ru.PosedMesh(body["world_verts"], body["faces"], body["Ts"], "cuda")
    assert pts.shape[0] * pts.shape[1] >= ru.SORT_QUERIES_FROM
    a = ru.warp_samples_to_canonical(pts, None, None, None, 0.05, mesh=mesh, return_query=True)
    old, ru.SORT_QUERIES_FROM = ru.SORT_QUERIES_FROM, 1 << 40
    try:
        b = ru.warp_samples_to_canonical(pts, None, None, None, 0.05, mesh=mesh, return_query=True)
    finally:
        ru.SORT_QUERIES_FROM = old
    for x, y in zip(a, b):
        assert torch.equal(x, y) or (torch.isnan(x) == torch.isnan(y)).all() and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
    assert 0.2 < float(a[3].float().mean()) < 1.0


def test_mesh_guided_near_far_against_oracle(body):
    from avatarcraft_b200.utils.ray_utils import geometry_guided_near_far
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 64, 64)
    gn, gf = wo.geometry_guided_near_far(o, d, body["world_verts"], 0.05)
    cn, cf = OracleNSR.near_far(o, d, 1.6)
    near_o = torch.where(torch.isinf(gn), cn[:, 0], gn)
    far_o = torch.where(torch.isinf(gf), cf[:, 0], gf)
    near, far = geometry_guided_near_far(o.cuda(), d.cuda(), body["world_verts"], 0.05, bound=1.6)
    hit = ~torch.isinf(gn)
    assert 0.05 < float(hit.float().mean()) < 0.9
    np.testing.assert_allclose(near.cpu().numpy(), near_o.numpy(), atol=2e-4)       # sqrt of a difference near grazing
    np.testing.assert_allclose(far.cpu().numpy(), far_o.numpy(), atol=2e-4)


def test_warped_render_against_oracle(body):
    """render_can=False end to end (32+32 samples, the reference's animate settings, render_warp.py:88-106)."""
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 64, 64)
    sel = torch.arange(64 * 26 + 14, 64 * 38, 19)                         # ~40 rays through the torso
    o, d = o[sel].contiguous(), d[sel].contiguous()
    ref = OracleNSR(sd).run(o, d, 32, 1.6, 32, verts=body["world_verts"], faces=body["faces"], Ts=body["Ts"])
    out = net.run(o.cuda()[None], d.cuda()[None], 32, 1.6, 32, None, 1.0, 0.0, render_can=False, verts=body["world_verts"],
                  faces=body["faces"], Ts=body["Ts"])
    torch.cuda.synchronize()
    rgb, rgb_o = out[3].reshape(-1, 3).cpu().numpy(), ref[3].reshape(-1, 3).numpy()
    assert psnr(rgb, rgb_o) >= 40.0
    dz = (out[9].cpu() - ref[9]).abs().max(1)[0].numpy()
    assert (dz <= 1e-4).mean() >= 0.8
    ok = dz <= 1e-5
    np.testing.assert_allclose(out[2].reshape(-1).cpu().numpy()[ok], ref[2].reshape(-1).numpy()[ok], atol=3e-3)
    hit = ref[2].reshape(-1).numpy() > 0.5
    assert hit.sum() >= 5                                                  # the warp actually lands samples inside the canonical body


def test_warp_skip_masked_leaves_the_image_bit_identical(body):
    """NeRFNetwork.warp_skip_masked (what render_warp.py turns on): the section-point warp stops at the mask distance and
    fully masked sample blocks are not evaluated.  Everything that reaches the image must be bit-identical to the exact path;
    inside the mask distance the bounded search must return the unbounded search's canonical points."""
    from avatarcraft_b200.utils.ray_utils import PosedMesh, warp_samples_to_canonical
    from avatarcraft_b200.utils.constant import DEFAULT_GEO_THRESH
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 64, 64)
    o, d = o.cuda().contiguous(), d.cuda().contiguous()
    kw = dict(render_can=False, verts=body["world_verts"], faces=body["faces"], Ts=body["Ts"])
    exact = net.run(o[None], d[None], 32, 1.6, 32, None, 1.0, 0.0, **kw)
    net.warp_skip_masked = True
    try:
        fast = net.run(o[None], d[None], 32, 1.6, 32, None, 1.0, 0.0, **kw)
    finally:
        net.warp_skip_masked = False
    torch.cuda.synchronize()
    for i, name in ((0, "depth"), (1, "weights"), (2, "weights_sum"), (3, "image"), (4, "normal_map"), (8, "alpha"), (9, "z_vals")):
        assert torch.equal(exact[i], fast[i]), name
    wsum = exact[2].reshape(-1)
    assert int((wsum > 0.5).sum()) >= 200 and int((wsum == 0).sum()) >= 1000          # body rays and empty rays both present
    # the bounded search itself: mask identical everywhere, canonical points identical wherever the mask is set
    mesh = PosedMesh(body["world_verts"], body["faces"], body["Ts"], "cuda")
    g = torch.Generator().manual_seed(3)
    v = torch.as_tensor(body["world_verts"], dtype=torch.float32)
    pts = torch.cat([v[torch.randint(0, v.shape[0], (20000,), generator=g)] + 0.15 * torch.randn(20000, 3, generator=g),
                     torch.rand(20000, 3, generator=g) * 3.2 - 1.6]).cuda().reshape(40, 1000, 3)
    can_a, mask_a = warp_samples_to_canonical(pts, None, None, None, DEFAULT_GEO_THRESH, mesh=mesh, product=True)
    can_b, mask_b = warp_samples_to_canonical(pts, None, None, None, DEFAULT_GEO_THRESH, mesh=mesh, product=True, masked_only=True)
    torch.cuda.synchronize()
    assert torch.equal(mask_a, mask_b)
    on = mask_a > 0.5
    assert 5000 < int(on.sum()) < 35000
    assert torch.equal(can_a[on], can_b[on])
    assert torch.equal(can_b[~on], pts[~on])


def test_render_driver_passes_warp_arguments(body):
    from avatarcraft_b200.utils.render_utils import render_instantnsr_naive
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 32, 32)
    rgb, eik, extra = render_instantnsr_naive(net, o.cuda(), d.cuda(), rays_per_batch=512, render_can=False, perturb=False,
                                              return_raw=True, verts=body["world_verts"], faces=body["faces"], Ts=body["Ts"],
                                              num_steps=32, upsample_steps=32, bound=1.6)
    assert rgb.shape == (1024, 3) and torch.isfinite(rgb).all() and extra["weight_sum"].shape == (1024, 1)


def test_entry_scripts_run_end_to_end(tmp_path):
    """render_canonical.py / render_warp.py / stylize.py in --synthetic mode at tiny sizes (assets of the reference are
    not redistributable): they must run, write their outputs, and stylize must change the parameters."""
    import os, subprocess, sys
    from tests.util import ROOT
    env = dict(os.environ, PYTHONPATH=ROOT)
    run = lambda *a: subprocess.run([sys.executable, *a], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    r = run(os.path.join(ROOT, "render_canonical.py"), "--synthetic", "--exp_name", "t", "--render_h", "32", "--render_w", "32", "--n_views", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(tmp_path / "demo" / "canonical_360" / "t" / "t_body.gif")
    r = run(os.path.join(ROOT, "render_warp.py"), "--synthetic", "--exp_name", "w", "--resolution", "32", "--max_frames", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(tmp_path / "demo" / "test_views" / "w" / "w_0001.png")
    r = run(os.path.join(ROOT, "stylize.py"), "--synthetic", "--exp_name", "s", "--render_h", "64", "--render_w", "64", "--n_views", "2",
            "--coarse_epochs", "1", "--fine_epochs", "0", "--subsample_scale", "2", "--batch_size", "512")
    assert r.returncode == 0, r.stderr[-2000:]
    sd = torch.load(tmp_path / "style" / "canonical_360" / "s" / "s.pth.tar", map_location="cpu")
    ref = state_dict("trained", 43)
    assert set(sd.keys()) == set(ref.keys())
    assert float((sd["encoder.embeddings"] - ref["encoder.embeddings"]).abs().max()) > 0


def test_stylize_with_sds_guidance_and_resume(tmp_path):
    """stylize.py --guidance sds (SD-1.5-shaped networks, random weights: the HF checkpoints are not available offline)
    for one view, then a second run resumed from its checkpoint: optimiser moments and the step counter come back."""
    import os, subprocess, sys
    from tests.util import ROOT
    env = dict(os.environ, PYTHONPATH=ROOT)
    run = lambda *a: subprocess.run([sys.executable, os.path.join(ROOT, "stylize.py"), "--synthetic", "--exp_name", "g", "--render_h", "64",
                                     "--render_w", "64", "--n_views", "1", "--fine_epochs", "0", "--subsample_scale", "1", "--batch_size", "4096",
                                     "--guidance", "sds", *a], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    r = run("--coarse_epochs", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    ck = tmp_path / "style" / "canonical_360" / "g" / "g.pth.tar"
    res = torch.load(str(ck)[:-len(".pth.tar")] + ".resume.pt", map_location="cpu", weights_only=False)
    assert res["step"] == 1 and res["optimizer"]["step"] == 1 and float(res["optimizer"]["exp_avg"].abs().max()) > 0
    first = torch.load(ck, map_location="cpu")
    r = run("--coarse_epochs", "2", "--resume", str(ck))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "resumed from" in r.stdout and "optimizer state restored" in r.stdout
    res2 = torch.load(str(ck)[:-len(".pth.tar")] + ".resume.pt", map_location="cpu", weights_only=False)
    assert res2["step"] == 2 and res2["optimizer"]["step"] == 2
    second = torch.load(ck, map_location="cpu")
    assert float((second["encoder.embeddings"] - first["encoder.embeddings"]).abs().max()) > 0


def test_stylize_resume_is_step_accurate_inside_an_epoch(tmp_path):
    """A periodic checkpoint taken in the middle of an epoch records the position inside the epoch's view permutation: the
    resumed run trains the remaining views only (same total step count as the uninterrupted run, reference file names
    <exp>_<step+1, 0-based: 4 digits>.pth.tar) and ends at (nearly) the same weights -- the hash-table gradient is summed
    with atomics, so not bit for bit."""
    import os, subprocess, sys
    from tests.util import ROOT
    env = dict(os.environ, PYTHONPATH=ROOT)
    def run(cwd, *a):
        os.makedirs(cwd, exist_ok=True)
        return subprocess.run([sys.executable, os.path.join(ROOT, "stylize.py"), "--synthetic", "--exp_name", "r", "--render_h", "32", "--render_w", "32",
                               "--n_views", "3", "--coarse_epochs", "1", "--fine_epochs", "0", "--subsample_scale", "1", "--batch_size", "1024",
                               "--guidance", "target", "--i_save", "2", *a], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    a, b = tmp_path / "a", tmp_path / "b"
    r = run(a)
    assert r.returncode == 0, r.stderr[-2000:]
    out_a = a / "style" / "canonical_360" / "r"
    assert (out_a / "r_0002.pth.tar").exists() and (out_a / "r_0004.pth.tar").exists()          # periodic (step 2) and final (3 + 1)
    mid = torch.load(out_a / "r_0002.resume.pt", map_location="cpu", weights_only=False)
    assert mid["step"] == 2 and mid["epoch"] == 0 and mid["extra"]["view_pos"] == 2
    r = run(b, "--resume", str(out_a / "r_0002.pth.tar"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "view 2" in r.stdout
    out_b = b / "style" / "canonical_360" / "r"
    end_b = torch.load(out_b / "r.resume.pt", map_location="cpu", weights_only=False)
    assert end_b["step"] == 3 and end_b["optimizer"]["step"] == 3                                 # one more step, not a replayed epoch
    wa, wb = torch.load(out_a / "r.pth.tar", map_location="cpu"), torch.load(out_b / "r.pth.tar", map_location="cpu")
    for k in wa:
        assert float((wa[k].float() - wb[k].float()).abs().max()) <= 2e-3, k
    # constant learning rate by default: the reference never steps its StepLR (stylize.py:214)
    assert end_b["optimizer"]["lr"] == 5e-3


def test_ray_coherent_warp_equals_per_point_search(body):
    """ac_warp_samples_to_canonical_rays (a thread walks 8 consecutive samples of a ray, each search seeded by the previous
    result) must reproduce the per-point search bit for bit: same faces, distances, closest and canonical points, mask."""
    from avatarcraft_b200.utils import ray_utils as ru
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 96, 96)
    z = torch.linspace(0.6, 2.8, 44)                              # 44 samples: the last chunk of 8 is ragged
    pts = (o[:, None, :] + d[:, None, :] * z[None, :, None]).cuda().contiguous()
    mesh = ru.PosedMesh(body["world_verts"], body["faces"], body["Ts"], "cuda")
    b = ru.warp_samples_to_canonical(pts, None, None, None, 0.05, mesh=mesh, return_query=True)
    old, ru.RAY_COHERENT = ru.RAY_COHERENT, True
    try:
        a = ru.warp_samples_to_canonical(pts, None, None, None, 0.05, mesh=mesh, return_query=True)
    finally:
        ru.RAY_COHERENT = old
    for name, x, y in zip(("can", "dirs", "closest", "mask", "face", "dist2"), a, b):
        same = torch.equal(x, y) or ((torch.isnan(x) == torch.isnan(y)).all() and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y)))
        assert same, name
    assert 0.02 < float(a[3].float().mean()) < 0.9


def test_warped_render_against_oracle_4096_rays(body):
    """render_can=False on a whole 64x64 frame (4096 rays, 32+32 samples = 393 216 closest-point queries against 13 776
    triangles): the oracle's exhaustive float64 closest-point scan runs in torch on the GPU (oracle/warp_oracle.py), everything
    else of the oracle on the CPU as usual."""
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0), 64, 64)
    orc = OracleNSR(sd)
    orc.warp_device = "cuda"
    ref = orc.run(o, d, 32, 1.6, 32, verts=body["world_verts"], faces=body["faces"], Ts=body["Ts"])
    out = net.run(o.cuda()[None], d.cuda()[None], 32, 1.6, 32, None, 1.0, 0.0, render_can=False, verts=body["world_verts"],
                  faces=body["faces"], Ts=body["Ts"])
    torch.cuda.synchronize()
    rgb, rgb_o = out[3].reshape(-1, 3).cpu().numpy(), ref[3].reshape(-1, 3).numpy()
    dz = (out[9].cpu() - ref[9]).abs().max(1)[0].numpy()
    hit = ref[2].reshape(-1).numpy() > 0.5
    print(f"warped render, 4096 rays: PSNR {psnr(rgb, rgb_o):.1f} dB, depth coincidence (1e-4) {(dz <= 1e-4).mean():.4f}, "
          f"{int(hit.sum())} rays hit the canonical body")
    assert psnr(rgb, rgb_o) >= 40.0
    assert (dz <= 1e-4).mean() >= 0.9
    ok = dz <= 1e-5
    np.testing.assert_allclose(out[2].reshape(-1).cpu().numpy()[ok], ref[2].reshape(-1).numpy()[ok], atol=3e-3)
    assert hit.sum() >= 200
