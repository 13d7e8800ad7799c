"""The occupancy-grid raymarching operators against the REFERENCE'S OWN kernels.

oracle/_ref/_ref_raymarching.so = the reference's unmodified raymarching/src/{raymarching.cu,bindings.cpp}
built for sm_100a by oracle/build_ref.py (test infrastructure only).  Sample slabs / alive slots are handed
out by atomics in arrival order in both implementations, so everything is compared per ray id."""
import importlib.machinery
import importlib.util
import os

import pytest
import torch

from tests.util import ROOT

pytestmark = pytest.mark.gpu
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_ref_raymarching.so")
BOUND, H = 1.0, 64


def ref_module():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/_ref_raymarching.so not built (run oracle/build_ref.py in the build container)")
    loader = importlib.machinery.ExtensionFileLoader("_ref_raymarching", REF_SO)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader("_ref_raymarching", loader))
    loader.exec_module(mod)
    return mod


def scene(n_rays, bound=BOUND, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1) * 2.2 * bound
    target = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2 * bound
    d = torch.nn.functional.normalize(target - o, dim=-1)
    d[:4] = torch.tensor([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.6, 0.8, 0.0]])        # axis-aligned: 1/0 = inf
    o[:4] = -2.0 * bound * d[:4] + torch.tensor([0.01, 0.02, 0.03])
    d[4] = torch.tensor([0.0, 0.0, 1.0]); o[4] = torch.tensor([5.0, 5.0, -3.0])                         # misses the box
    ax = (torch.arange(H) + 0.5) / H * 2 - 1
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = torch.sqrt(X * X + 1.4 * Y * Y + Z * Z)
    grid = torch.where((r < 0.55) | ((X - 0.5).abs() + (Y + 0.4).abs() + Z.abs() < 0.25), 30.0, 0.0).contiguous()
    return o.cuda().contiguous(), d.cuda().contiguous(), grid.cuda()


def per_ray(rays, *arrays):
    """Re-order the sample arrays by ray id: returns (counts[N], concatenated samples ...)."""
    rays = rays.long()
    order = torch.argsort(rays[:, 0])
    rid, off, cnt = rays[order, 0], rays[order, 1], rays[order, 2]
    assert torch.equal(rid, torch.arange(rays.shape[0], device=rays.device))
    start = torch.cumsum(cnt, 0) - cnt
    idx = torch.repeat_interleave(off - start, cnt) + torch.arange(int(cnt.sum()), device=rays.device)
    return (cnt,) + tuple(a[idx] for a in arrays)


def fake_field(xyzs, dirs):
    """A deterministic stand-in for the network: alpha, rgb, normal from the sample position."""
    alpha = torch.sigmoid(8.0 * (0.5 - xyzs.norm(dim=-1))) * 0.35
    rgb = torch.sigmoid(3.0 * xyzs + dirs)
    nrm = torch.nn.functional.normalize(xyzs + 1e-3, dim=-1)
    return alpha.contiguous(), rgb.contiguous(), nrm.contiguous()


@pytest.mark.parametrize("bound,perturb", [(1.0, False), (1.0, True), (2.0, False)])
def test_march_rays_train_matches_reference_kernel(bound, perturb):
    from avatarcraft_b200 import raymarching as rm
    ref = ref_module()
    N = 20000
    o, d, grid = scene(N, bound)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, bound, grid, 12.0, 0, perturb=perturb, force_all_rays=True)
    M = N * 1024
    rx, rd_, rdl = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, device="cuda")
    rrays, rcnt = torch.empty(N, 3, dtype=torch.int32, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda")
    ref.march_rays_train(o, d, grid, 12.0, 0, bound, N, H, M, rx, rd_, rdl, rrays, rcnt, int(perturb))
    torch.cuda.synchronize()
    assert int(rcnt[1]) == N and int(rcnt[0]) == xyzs.shape[0] > 10 * N
    c_m, x_m, d_m, t_m = per_ray(rays, xyzs, dirs, deltas)
    c_r, x_r, d_r, t_r = per_ray(rrays, rx, rd_, rdl)
    assert torch.equal(c_m, c_r), f"{int((c_m != c_r).sum())} rays differ in step count"
    assert int(c_m[4]) == 0 and int(c_m.max()) <= 1024
    assert torch.equal(x_m, x_r) and torch.equal(d_m, d_r) and torch.equal(t_m, t_r)


def test_composite_rays_train_forward_backward_match_reference_kernel():
    from avatarcraft_b200 import raymarching as rm
    ref = ref_module()
    N = 8192
    o, d, grid = scene(N)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, BOUND, grid, 12.0, 0, force_all_rays=True)
    M = xyzs.shape[0]
    alpha, rgb, _ = fake_field(xyzs, dirs)
    alpha[: M // 3] = (alpha[: M // 3] * 2.5).clamp(max=0.95)             # some rays hit the T < 1e-4 early-out
    a, c = alpha.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    ws, img = rm.composite_rays_train(a, c, deltas, rays, BOUND)
    g_ws, g_img = torch.randn_like(ws), torch.randn_like(img)
    (ws * g_ws).sum().add((img * g_img).sum()).backward()
    ws_r, img_r = torch.empty(N, device="cuda"), torch.empty(N, 3, device="cuda")
    ref.composite_rays_train_forward(alpha, rgb, deltas, rays, BOUND, M, N, ws_r, img_r)
    ga_r, gc_r = torch.zeros_like(alpha), torch.zeros_like(rgb)
    ref.composite_rays_train_backward(g_ws, g_img, alpha, rgb, deltas, rays, ws_r, img_r, BOUND, M, N, ga_r, gc_r)
    torch.cuda.synchronize()
    assert float(ws_r.max()) > 0.99 and float(ws_r.min()) == 0.0
    assert torch.equal(ws.detach(), ws_r) and torch.equal(img.detach(), img_r)
    assert torch.equal(c.grad, gc_r)
    assert torch.allclose(a.grad, ga_r, rtol=0, atol=2e-7 * float(ga_r.abs().max()))      # fma contraction choices


def test_march_rays_perturbed_single_call_matches_reference_kernel():
    """The inference jitter is seeded by the ALIVE SLOT (raymarching.cu:543), so after the first atomic compaction it is
    not reproducible even between two runs of the reference; pin it on one call with an identical alive list."""
    from avatarcraft_b200 import raymarching as rm
    ref = ref_module()
    N, n_step = 16384, 8
    o, d, grid = scene(N)
    alive = torch.randperm(N, generator=torch.Generator().manual_seed(5)).int().cuda()[: N // 2].contiguous()
    ts = (0.05 + torch.rand(N // 2, generator=torch.Generator().manual_seed(6))).cuda()
    near, far = torch.full((N,), 0.05, device="cuda"), torch.full((N,), 4.0, device="cuda")
    x_m, d_m, t_m = rm.march_rays(N // 2, n_step, alive, ts, o, d, BOUND, grid, 12.0, near, far, perturb=7)
    M = N // 2 * n_step
    x_r, d_r, t_r = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, 2, device="cuda")
    ref.march_rays(N // 2, n_step, alive, ts, o, d, BOUND, H, grid, 12.0, near, far, x_r, d_r, t_r, 7)
    torch.cuda.synchronize()
    assert float((t_r[:, 0] > 0).float().mean()) > 0.2
    assert torch.equal(x_m, x_r) and torch.equal(d_m, d_r) and torch.equal(t_m, t_r)


@pytest.mark.parametrize("perturb", [0])
def test_inference_loop_matches_reference_kernels(perturb):
    """march_rays -> composite_rays -> compact_rays until every ray retires (the reference's run_cuda loop shape)."""
    from avatarcraft_b200 import raymarching as rm
    ref = ref_module()
    N = 16384
    o, d, grid = scene(N)
    near = torch.full((N,), 0.05, device="cuda")
    far = torch.full((N,), 4.0, device="cuda")

    def run(mine):
        w, dep = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
        img, nm = torch.zeros(N, 3, device="cuda"), torch.zeros(N, 3, device="cuda")
        alive = [torch.arange(N, dtype=torch.int32, device="cuda"), torch.zeros(N, dtype=torch.int32, device="cuda")]
        ts = [near.clone(), torch.zeros(N, device="cuda")]
        n_alive, it, n_step = N, 0, 16
        while n_alive > 0 and it < 200:
            cur, nxt = it % 2, (it + 1) % 2
            if mine:
                xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive[cur], ts[cur], o, d, BOUND, grid, 12.0, near, far, perturb=perturb)
            else:
                M = n_alive * n_step
                xyzs, dirs, deltas = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, 2, device="cuda")
                ref.march_rays(n_alive, n_step, alive[cur], ts[cur], o, d, BOUND, H, grid, 12.0, near, far, xyzs, dirs, deltas, perturb)
            a, c, nr = fake_field(xyzs, dirs)
            counter = torch.zeros(1, dtype=torch.int32, device="cuda")
            if mine:
                rm.composite_rays(n_alive, n_step, alive[cur], ts[cur], a, c, nr, deltas, w, dep, img, nm)
                rm.compact_rays(n_alive, alive[nxt], alive[cur], ts[nxt], ts[cur], counter)
            else:
                ref.composite_rays(n_alive, n_step, alive[cur], ts[cur], a, c, nr, deltas, w, dep, img, nm)
                ref.compact_rays(n_alive, alive[nxt], alive[cur], ts[nxt], ts[cur], counter)
            n_alive = int(counter.item())
            it += 1
        return w, dep, img, nm, it

    w_m, d_m, i_m, n_m, it_m = run(True)
    w_r, d_r, i_r, n_r, it_r = run(False)
    assert it_m == it_r and 2 < it_m < 200
    assert float(w_r.max()) > 0.9
    # the field is evaluated by torch on identical sample positions, so everything is bit-identical
    assert torch.equal(w_m, w_r) and torch.equal(d_m, d_r) and torch.equal(i_m, i_r) and torch.equal(n_m, n_r)


def test_raymarching_rejects_null_and_degenerate_arguments():
    from avatarcraft_b200 import _lib
    lib = _lib.lib()
    assert lib.ac_march_rays_train(None, None, None, 1.0, 0, 1.0, 4, 64, 4096, None, None, None, None, None, 0, None) != 0
    assert lib.ac_compact_rays(0, None, None, None, None, None, None) != 0
