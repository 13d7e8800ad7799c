"""world_size-2 gloo tests (CPU) of the N>1 host logic: patch sharding and the single flat
gradient all-reduce reproduce the single-process gradient of the reference's patch loop."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avatarcraft_b200.utils.distributed import allreduce_gradients, shard_patches, shard_rays


def test_shard_patches_partitions_every_ray_once():
    for n, bs, world in [(65536, 4096, 8), (65536, 4096, 3), (4096, 4096, 8), (4096, 4096, 2), (5000, 4096, 4), (100, 7, 2)]:
        seen = torch.zeros(n, dtype=torch.int32)
        for r in range(world):
            for s, e, scale in shard_patches(n, bs, r, world):
                seen[s:e] += 1
                patch_len = min(bs, n - (s // bs) * bs)
                assert 0 < scale <= 1.0 and abs(scale - (e - s) / patch_len) < 1e-12 or scale == 1.0
        assert int(seen.min()) == 1 and int(seen.max()) == 1, (n, bs, world)
    lo, hi = shard_rays(65536, 3, 8)
    assert (lo, hi) == (24576, 32768)


def _loss(model, x, start, end, scale):
    """Stand-in for one rendered patch: a sum-type term and a mean-type term (like eikonal / opacity)."""
    y = model(x[start:end])
    return (y * torch.arange(start, end, dtype=torch.float32)[:, None]).sum() + scale * 7.0 * (y ** 2).mean()


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, n, bs, ref_grads, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.Softplus(beta=100), torch.nn.Linear(16, 2))
    x = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    for s, e, scale in shard_patches(n, bs, rank, world):
        _loss(model, x, s, e, scale).backward()
    reduced = allreduce_gradients(model.parameters())
    err = max(float((p.grad - g).abs().max() / (g.abs().max() + 1e-12)) for p, g in zip(model.parameters(), ref_grads))
    q.put((rank, reduced, err))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,bs", [(4096, 512), (600, 600)])
def test_two_rank_patch_step_equals_single_process(n, bs):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.Softplus(beta=100), torch.nn.Linear(16, 2))
    x = torch.rand(n, 3, generator=torch.Generator().manual_seed(1))
    for s in range(0, n, bs):
        _loss(model, x, s, min(s + bs, n), 1.0).backward()
    ref = [p.grad.clone() for p in model.parameters()]
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, bs, ref, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, reduced, err in res:
        assert reduced == sum(p.numel() for p in model.parameters())
        # relative to the largest entry; (600, 600) = ONE patch split over both ranks with mean_scale 1/2 each
        assert err < 1e-5, (rank, err)


def _cfg_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from avatarcraft_b200.models import diffusion, sd_vae
    from tests.test_sds_cpu import _StubUNet
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    sd = diffusion.StableDiffusion("cpu", "1.5", unet=_StubUNet(), vae=sd_vae.AutoencoderKL.tiny(), text_encoder=diffusion.HashTextEncoder(32))
    assert sd.cfg_parallel is False                 # off unless the caller guarantees identical (t, noise) on every rank
    sd.cfg_parallel = True                          # ... which the shared per-step seed does (stylize.py)
    emb = torch.stack([torch.full((77, 32), 0.2), torch.full((77, 32), 0.5)])
    rgb = torch.rand(16 * 16, 3, generator=torch.Generator().manual_seed(5))
    try:
        sd.pixel_gradient(emb, rgb, 16, 16, guidance_scale=100)           # no seed: the ranks would draw different noise
        refused = False
    except RuntimeError:
        refused = True
    assert refused
    g = sd.pixel_gradient(emb, rgb, 16, 16, guidance_scale=100, seed=11)
    q.put((rank, g.numpy()))              # by value: a tensor's shared-memory handle dies with this process
    dist.destroy_process_group()


def test_cfg_pair_split_over_two_ranks_equals_single_process():
    """the SD step does not shard by rays; the (uncond, text) UNet pair does.  Two gloo ranks, each
    evaluating one half + one all-gather, must reproduce the single-process pixel gradient bit for bit."""
    import torch
    import torch.multiprocessing as mp
    from avatarcraft_b200.models import diffusion, sd_vae
    from tests.test_sds_cpu import _StubUNet
    torch.manual_seed(0)
    sd = diffusion.StableDiffusion("cpu", "1.5", unet=_StubUNet(), vae=sd_vae.AutoencoderKL.tiny(), text_encoder=diffusion.HashTextEncoder(32))
    emb = torch.stack([torch.full((77, 32), 0.2), torch.full((77, 32), 0.5)])
    rgb = torch.rand(16 * 16, 3, generator=torch.Generator().manual_seed(5))
    want = sd.pixel_gradient(emb, rgb, 16, 16, guidance_scale=100, seed=11)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_cfg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = {r: torch.from_numpy(g) for r, g in (q.get(timeout=300) for _ in range(2))}
    for p in ps:
        p.join(timeout=60)
    assert torch.equal(got[0], got[1]) and torch.allclose(got[0], want, atol=1e-7) and float(want.abs().max()) > 0


def _shard_render_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from avatarcraft_b200.utils.distributed import render_rays_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = torch.arange(24.0).reshape(8, 3); d = torch.ones(8, 3)
    calls = []

    def fake_render(a, b):                      # a "render" whose pixel depends only on its own ray
        calls.append(a.shape[0])
        return a * 2.0 + b

    full = render_rays_sharded(fake_render, o, d, rank, world)
    odd = render_rays_sharded(fake_render, o[:7], d[:7], rank, world)        # 7 rays on 2 ranks: padded, gathered, trimmed
    assert torch.equal(odd, o[:7] * 2.0 + 1.0)
    q.put((rank, full.numpy(), calls[:1]))
    dist.destroy_process_group()


def test_pass1_ray_sharding_reassembles_the_whole_image():
    """Pass 1 of a multi-GPU step: each rank renders n/world rays, ONE all-gather gives every rank the full image."""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_shard_render_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = {r: (torch.from_numpy(img), calls) for r, img, calls in (q.get(timeout=300) for _ in range(2))}
    for p in ps:
        p.join(timeout=60)
    want = torch.arange(24.0).reshape(8, 3) * 2.0 + 1.0
    for r in range(2):
        assert torch.equal(got[r][0], want) and got[r][1] == [4]       # half of the rays rendered locally, whole image held


def _masked_mean_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from avatarcraft_b200.utils.distributed import masked_mean_share, shard_patches
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(3)
    w = torch.rand(600, 8, generator=gen).requires_grad_(True)                 # stands in for the parameters
    pts = torch.rand(600, 8, generator=gen) * 2.0                               # |p| decides the eikonal mask
    (s, e, scale), = shard_patches(600, 600, rank, world)                       # ONE patch split over both ranks
    assert scale == 0.5
    relax = (pts[s:e] < 1.2).float()
    err = (w[s:e] * 3.0 - 1.0) ** 2
    eik = (relax * err).sum() / (relax.sum() + 1e-5)                            # this rank's masked mean (models/instant_nsr.py:266-272)
    (eik * masked_mean_share(relax.sum())).backward()
    g = w.grad.clone()
    dist.all_reduce(g)
    q.put((rank, g.numpy(), float(relax.sum())))
    dist.destroy_process_group()


def test_split_patch_masked_eikonal_equals_single_process():
    """the eikonal term is a masked mean whose mask count differs per rank; weighting each rank by its ray
    share (1/2) is wrong, weighting by its share of the mask count reproduces the single-process gradient."""
    import torch
    import torch.multiprocessing as mp
    gen = torch.Generator().manual_seed(3)
    w = torch.rand(600, 8, generator=gen).requires_grad_(True)
    pts = torch.rand(600, 8, generator=gen) * 2.0
    relax = (pts < 1.2).float()
    ((relax * (w * 3.0 - 1.0) ** 2).sum() / (relax.sum() + 1e-5)).backward()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_masked_mean_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = {r: (torch.from_numpy(g), c) for r, g, c in (q.get(timeout=300) for _ in range(2))}
    for p in ps:
        p.join(timeout=60)
    assert got[0][1] != got[1][1]                                               # the two halves do have different mask counts
    for r in range(2):
        assert float((got[r][0] - w.grad).abs().max()) <= 1e-6 * float(w.grad.abs().max())
