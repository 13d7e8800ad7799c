"""GPU suite for the stages either side of the render core: on-device ray generation and
backgrounds, the flat Adam update, the SDF lattice for mesh extraction.  Checkers are the host restatements the
CPU suite pins to the reference (utils/synthetic.pinhole_rays, utils/ray_gen.gen_rays_pose), torch.optim.Adam
(the optimizer the reference uses, stylize.py:355-363) and plain torch for the image-space ops."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from avatarcraft_b200.utils import synthetic as syn
from avatarcraft_b200.utils import ray_gen
from avatarcraft_b200.utils.render_utils import select_background_device
from tests.util import state_dict, gpu_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wh", [(64, 64), (256, 256), (96, 40)])
def test_pinhole_rays_match_host_restatement(wh):
    W, H = wh
    for angle in (-180.0, 30.0, 77.5):
        c2w = syn.orbit_pose(angle, dist=1.7, center=(0.0, 0.1, 0.0))
        o_h, d_h = syn.pinhole_rays(c2w, W, H)
        o, d = ray_gen.pinhole_rays_device(c2w, W, H, "cuda")
        assert torch.equal(o.cpu(), o_h)                                   # float32(camera centre), bit-exact
        # numpy's matmul may contract differently in float64; after rounding to float32 at most 1 ulp remains
        assert float((d.cpu() - d_h).abs().max()) <= 2.4e-7
        assert float((d.norm(dim=-1) - 1).abs().max()) < 1e-6


def test_gen_rays_pose_matches_host_restatement():
    K = ray_gen.dataset_intrinsics(512, 512)
    pose = torch.from_numpy(syn.orbit_pose(12.0)).float()
    for level in (1, 2, 4):
        o_h, v_h = ray_gen.gen_rays_pose(pose, K, 512, 512, level)
        o, v = ray_gen.gen_rays_pose_device(pose, K, 512, 512, level, "cuda")
        assert o.shape == v_h.shape and torch.equal(o.cpu(), o_h.contiguous())
        assert float((v.cpu() - v_h).abs().max()) <= 5e-7


def test_backgrounds():
    n = 4096
    assert torch.equal(select_background_device(n, 0, "cuda"), torch.ones(n, 3, device="cuda"))
    assert torch.equal(select_background_device(n, 1, "cuda"), torch.zeros(n, 3, device="cuda"))
    big = select_background_device(1 << 20, 2, "cuda", seed=7)
    assert torch.equal(big[:, 0], big[:, 1]) and torch.equal(big[:, 0], big[:, 2])      # grey
    assert 0.0 <= float(big.min()) and float(big.max()) <= 1.0
    assert abs(float(big[:, 0].mean()) - 0.5) < 1e-3 and abs(float(big[:, 0].std()) - 0.1) < 1e-3
    assert torch.equal(big, select_background_device(1 << 20, 2, "cuda", seed=7))        # counter-based: reproducible
    assert not torch.equal(big, select_background_device(1 << 20, 2, "cuda", seed=8))
    # chessboard: torchvision GaussianBlur((5, 9), sigma) on the 0.8 / 0.2 board, reflect padding
    side, sigma = 64, 1.3
    dev = select_background_device(side * side, 3, "cuda", sigma=sigma).cpu()
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    cell = side // 10
    board = torch.from_numpy(np.where(((ii // cell) + (jj // cell)) % 2 == 0, 0.8, 0.2).astype(np.float32))

    def taps(k):
        x = torch.linspace(-(k - 1) * 0.5, (k - 1) * 0.5, k)
        p = torch.exp(-0.5 * (x / sigma) ** 2)
        return p / p.sum()
    k2 = taps(9)[:, None] @ taps(5)[None, :]                                            # [ky=9, kx=5]
    ref = F.conv2d(F.pad(board[None, None], (2, 2, 4, 4), mode="reflect"), k2[None, None])[0, 0]
    assert float((dev[:, 0].reshape(side, side) - ref).abs().max()) < 2e-6
    with pytest.raises(RuntimeError):
        select_background_device(1000, 3, "cuda", sigma=1.0)                              # not a square batch


def test_flat_adam_matches_torch_adam():
    from avatarcraft_b200.utils.optim import FlatAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(1001, 2), (64, 35), (64,), (16, 64), (1,), (3, 64)]
    ps_a = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    ref = torch.optim.Adam(ps_b, lr=5e-3)
    opt = FlatAdam(ps_a, lr=5e-3)
    for step in range(6):
        opt.zero_grad(); ref.zero_grad()
        for pa, pb in zip(ps_a, ps_b):
            gr = torch.randn(pa.shape, generator=g).cuda()
            if pa.dim() == 2 and pa.shape[0] > 100:
                gr[::3] = 0.0                                                # slots without gradient (hash-table style)
                if step < 2:
                    gr[1::3] = 0.0                                           # ... some only get one later
            (pa * gr).sum().backward()
            (pb * gr).sum().backward()
        opt.step(); ref.step()
        for pa, pb in zip(ps_a, ps_b):
            assert float((pa - pb).abs().max()) <= 2e-7 + 1e-6 * float(pb.abs().max()), step
    # version counters were bumped (the packed-MLP cache keys on them)
    assert all(p._version > 0 for p in ps_a)
    st = ref.state[ps_b[0]]
    a0, n0 = opt._spans[0]
    assert float((opt.exp_avg[a0:a0 + n0].view_as(ps_a[0]) - st["exp_avg"]).abs().max()) < 1e-6
    assert float((opt.exp_avg_sq[a0:a0 + n0].view_as(ps_a[0]) - st["exp_avg_sq"]).abs().max()) < 1e-6


def test_flat_adam_on_the_model_repacks_weights():
    """After an update through raw pointers the render must see the new weights (blob cache invalidated) and the
    state-dict must still be the reference layout."""
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    from avatarcraft_b200.utils.optim import FlatAdam
    sd = state_dict("trained", 43)
    net = gpu_model(sd, train=True)
    x = (torch.rand(4096, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    before = net.forward_sdf(x, 1.6).clone()
    opt = FlatAdam(net.parameters(), lr=1e-2)
    opt.zero_grad()
    for p in net.parameters():
        p.grad.copy_(torch.randn(p.shape, generator=torch.Generator().manual_seed(2)).cuda())
    opt.step()
    after = net.forward_sdf(x, 1.6)
    assert float((after - before).abs().max()) > 1e-4
    fresh = NeRFNetwork()
    fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()})
    fresh = fresh.cuda().eval()
    assert torch.equal(fresh.forward_sdf(x, 1.6), after)
    assert set(net.state_dict().keys()) == set(sd.keys())


def test_sdf_grid_matches_pointwise_queries():
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    res, bound = 48, 1.6
    u = net.extract_fields(bound, res)
    assert u.shape == (res, res, res)
    X = torch.linspace(-bound, bound, res)
    xx, yy, zz = torch.meshgrid(X, X, X, indexing="ij")
    pts = torch.stack([xx, yy, zz], -1).reshape(-1, 3).cuda()
    ref = net.forward_sdf(pts, bound)[:, 0].reshape(res, res, res)
    assert float((u - ref).abs().max()) < 1e-5


def test_extract_geometry_is_a_closed_surface_on_the_zero_set():
    sd = state_dict("trained", 43)
    net = gpu_model(sd)
    verts, tris = net.extract_geometry(1.6, 96)
    assert verts.shape[0] > 1000 and tris.shape[0] > 2000
    # every vertex lies on the zero set (up to the trilinear interpolation error of a 3.2/95 lattice)
    s = net.forward_sdf(torch.from_numpy(verts).cuda(), 1.6)[:, 0]
    assert float(s.abs().max()) < 5e-2 and float(s.abs().mean()) < 3e-3      # bumpy synthetic field: hash detail below the lattice pitch
    # watertight: every undirected edge is shared by exactly two triangles
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    und = np.sort(e, 1)
    _, counts = np.unique(und, axis=0, return_counts=True)
    assert (counts == 2).all()
    # outward orientation: the signed volume is positive and close to a radius-0.5 ball's
    p0, p1, p2 = verts[tris[:, 0]].astype(np.float64), verts[tris[:, 1]].astype(np.float64), verts[tris[:, 2]].astype(np.float64)
    vol = float(np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6.0)
    assert 0.2 < vol < 1.2, vol


def test_error_codes_of_the_new_entry_points():
    """Invalid arguments come back as AC_E_INVALID_ARG (the shim raises RuntimeError), never as a launch."""
    import ctypes
    from avatarcraft_b200 import _lib
    L = _lib.lib()
    t = torch.zeros(64, device="cuda")
    n0 = L.ac_launch_count()
    assert L.ac_gen_rays(None, 1.0, 1.0, 0.0, 0.0, 4, 4, 0.0, 1.0, 0.0, 1.0, 0, _lib.ptr(t), _lib.ptr(t), None) == _lib.AC_E_INVALID_ARG
    m = np.eye(4)
    assert L.ac_gen_rays(m.ctypes.data_as(ctypes.c_void_p), 1.0, 1.0, 0.0, 0.0, 4, 4, 0.0, 1.0, 0.0, 1.0, 7, _lib.ptr(t), _lib.ptr(t), None) == _lib.AC_E_INVALID_ARG
    assert L.ac_select_background(0, 0, 0, 1.0, _lib.ptr(t), None) == _lib.AC_E_INVALID_ARG
    assert L.ac_adam_step(_lib.ptr(t), _lib.ptr(t), _lib.ptr(t), _lib.ptr(t), 64, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, None) == _lib.AC_E_INVALID_ARG   # step 0
    assert L.ac_adam_step(ctypes.c_void_p(t.data_ptr() + 4), _lib.ptr(t), _lib.ptr(t), _lib.ptr(t), 8, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == _lib.AC_E_INVALID_ARG   # misaligned
    assert L.ac_sd_gemm_f16(_lib.ptr(t), _lib.ptr(t), None, None, 0, None, _lib.ptr(t), 0, 4, 4, 4, 3, 8, 4, 0, 1, 1, 0, 0, 0, 0, 0, 0, None) == _lib.AC_E_INVALID_ARG   # lda % 8
    assert L.ac_sd_softmax_f16(_lib.ptr(t), 4, 8, 4, 8, 1.0, _lib.ptr(t), None) == _lib.AC_E_INVALID_ARG      # ld_in < L
    assert L.ac_launch_count() == n0
    with pytest.raises(RuntimeError):
        _lib.check(_lib.AC_E_INVALID_ARG, "x")


def test_iso_surface_against_lattice_edge_oracle_and_analytic_sphere():
    """ac_iso_surface (marching tetrahedra) against what pins ANY lattice iso-surface extractor, the reference's
    mcubes.marching_cubes included (models/instant_nsr.py:733-752): a vertex sits on every lattice edge whose end values straddle
    the threshold, at the linear-interpolation point.  The axis-aligned edges are shared by marching cubes and marching
    tetrahedra, so the numpy oracle below (all axis-edge crossings) must be a subset of the device's vertices, to fp32 rounding;
    and on an analytic sphere the surface must converge like h^2."""
    import ctypes
    from scipy.spatial import cKDTree
    from avatarcraft_b200 import _lib
    res, r0 = 64, 0.9
    lo, hi = np.array([-1.3, -1.2, -1.1], np.float32), np.array([1.3, 1.2, 1.1], np.float32)
    ax = [np.linspace(lo[i], hi[i], res, dtype=np.float64) for i in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    u = (np.sqrt(X * X + Y * Y + Z * Z) - r0).astype(np.float32)                     # signed distance of a sphere, [i,j,k]
    vol = torch.from_numpy(u).cuda().contiguous()
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    L = _lib.lib()

    def run(cap, pos, key):
        counter.zero_()
        _lib.check(L.ac_iso_surface(_lib.ptr(vol), lo.ctypes.data_as(ctypes.c_void_p), hi.ctypes.data_as(ctypes.c_void_p), res, 0.0,
                                    None if pos is None else _lib.ptr(pos), None if key is None else _lib.ptr(key), cap, _lib.ptr(counter),
                                    _lib.stream_ptr()), "ac_iso_surface")
        return int(counter.item())
    n = run(0, None, None)
    pos = torch.empty(n, 3, 3, device="cuda"); key = torch.empty(n, 3, dtype=torch.int64, device="cuda")
    assert run(n, pos, key) == n
    verts = pos.reshape(-1, 3).cpu().numpy().astype(np.float64)
    # oracle: crossings of the three families of axis-aligned lattice edges
    P = np.stack([X, Y, Z], -1)
    want = []
    for axis in range(3):
        a = [slice(None)] * 3; b = [slice(None)] * 3
        a[axis], b[axis] = slice(0, res - 1), slice(1, res)
        ua, ub = u[tuple(a)].astype(np.float64), u[tuple(b)].astype(np.float64)
        cross = (ua < 0.0) != (ub < 0.0)
        t = (0.0 - ua[cross]) / (ub[cross] - ua[cross])
        want.append(P[tuple(a)][cross] + t[:, None] * (P[tuple(b)][cross] - P[tuple(a)][cross]))
    want = np.concatenate(want)
    d, _ = cKDTree(verts).query(want)
    print(f"iso-surface: {n} triangles; {len(want)} axis-edge crossings of the oracle, max distance to a device vertex {d.max():.2e}")
    assert len(want) > 3000 and d.max() < 2e-6
    # analytic: every vertex within O(h^2) of the sphere, enclosed volume within 0.5 %
    h = float(((hi - lo) / (res - 1)).max())
    rad = np.linalg.norm(verts, axis=1)
    assert np.abs(rad - r0).max() < 0.5 * h * h / r0 + 1e-5
    p0, p1, p2 = (pos[:, i].cpu().numpy().astype(np.float64) for i in range(3))
    volume = abs(float(np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6.0))
    assert abs(volume / (4.0 / 3.0 * np.pi * r0 ** 3) - 1.0) < 5e-3
