"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY; runs in the BUILD container only.

Imports the reference's own `models/instant_nsr.py` from /root/reference (read-only, never
copied), slots oracle/hashgrid.py in for the CUDA-only `_backend`
(encoder/hashencoder/hashgrid.py:9,38 -- the kernel has no CPU path, hashencoder.cu:414-418),
runs the reference `NeRFRenderer.run` on seeded synthetic inputs, checks that the
restatement in oracle/nsr_oracle.py agrees with it, and writes small golden fixtures to
tests/golden/*.npz.  /root/reference does not exist on the GPU box, so tests read only the
committed fixtures.

    python -m oracle.make_golden
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hashgrid as ohg                      # noqa: E402
from oracle.nsr_oracle import OracleNSR                 # noqa: E402
from avatarcraft_b200.utils import synthetic as syn     # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for name in ("mcubes", "trimesh", "igl"):          # mesh export / plotting / warp only
        sys.modules.setdefault(name, types.ModuleType(name))
    be = types.SimpleNamespace(hash_encode_forward=ohg.hash_encode_forward,
                               hash_encode_backward=ohg.hash_encode_backward)
    for pkg in ("encoder.hashencoder.backend", "encoder.shencoder.backend"):
        m = types.ModuleType(pkg)
        m._backend = be
        sys.modules[pkg] = m
    sys.path.insert(0, REF)
    warnings.simplefilter("ignore")
    import models.instant_nsr as ref_nsr
    return ref_nsr


def ref_run(ref_nsr, sd, rays_o, rays_d, num_steps, upsample_steps, bound, training=False, seed=None, bg=None):
    net = ref_nsr.NeRFNetwork()
    net.load_state_dict(sd)
    net.train(training)
    if seed is not None:
        torch.manual_seed(seed)
    with torch.no_grad():
        return net.run(rays_o[None], rays_d[None], num_steps, bound, upsample_steps, bg,
                       cos_anneal_ratio=1.0, normal_epsilon_ratio=0.0, render_can=True,
                       perturb_overwrite=training)


def pack(out):
    depth, weights, wsum, image, nmap, eik, _, color, alpha, z = out
    return dict(depth=depth.reshape(-1).numpy(), weights=weights.numpy(), weight_sum=wsum.reshape(-1).numpy(),
                rgb=image.reshape(-1, 3).numpy(), normal=nmap.numpy(), eikonal=np.float32(float(eik)),
                pts_color=color.numpy(), pts_alpha=alpha.numpy(), z_vals=z.numpy())


def training_loss(out, pixel_grad):
    """The three terms stylize.py back-propagates per patch (:163-193): pixel gradient . rgb,
    w_eikonal * eikonal, 1e5 * smooth_l1(clamp(opacity), target) -- target 0.5 stands in for net_gt."""
    rgb, wsum, eik = out[3].reshape(-1, 3), out[2], out[5]
    opacity = torch.nn.functional.smooth_l1_loss(wsum.clamp(0.0, 1.0), torch.full_like(wsum, 0.5)) * 1e5
    return (rgb * pixel_grad).sum() + 0.01 * eik + opacity * 1e-3


def compare(a, b, tag):
    worst = 0.0
    for k in a:
        d = float(np.max(np.abs(np.asarray(a[k], dtype=np.float64) - np.asarray(b[k], dtype=np.float64))))
        worst = max(worst, d)
        print(f"  [{tag}] {k:12s} max|ref-oracle| = {d:.3e}")
    return worst


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_nsr = import_reference()
    cases = [
        # name, checkpoint kind, seed, image WxH, ray subsample stride, steps, upsample, training
        ("c1_init_64x64_16p16", "init", 42, 64, 16, 16, 16, False),
        ("c2_trained_256x256_64p64", "trained", 43, 256, 257, 64, 64, False),
        ("c4_trained_256x256_32p32", "trained", 43, 256, 509, 32, 32, False),
        ("c3_trained_jitter_64p64", "trained", 43, 256, 1021, 64, 64, True),
    ]
    for name, kind, seed, wh, stride, ns, us, training in cases:
        sd = syn.synthetic_state_dict(kind, seed)
        o, d = syn.pinhole_rays(syn.orbit_pose(30.0 if kind == "trained" else 0.0), wh, wh)
        sel = torch.arange(0, o.shape[0], stride)
        ro, rd = o[sel].contiguous(), d[sel].contiguous()
        jitter = None
        if training:
            torch.manual_seed(1234)
            jitter = torch.rand(ro.shape[0], ns)        # the reference's first rand draw (:162)
        bg = None
        out_ref = pack(ref_run(ref_nsr, sd, ro, rd, ns, us, 1.6, training, 1234 if training else None, bg))
        out_orc = pack(OracleNSR(sd).run(ro, rd, ns, 1.6, us, bg, jitter=jitter))
        worst = compare(out_ref, out_orc, name)
        assert worst < 2e-5, f"oracle restatement disagrees with the reference on {name}: {worst}"
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), rays_o=ro.numpy(), rays_d=rd.numpy(),
                            jitter=np.zeros(0, np.float32) if jitter is None else jitter.numpy(),
                            num_steps=ns, upsample_steps=us, bound=np.float32(1.6), kind=kind, seed=seed,
                            state_checksum=syn.state_checksum(sd), **out_ref)
        print(f"wrote {name}.npz  rays={ro.shape[0]} wsum mean={out_ref['weight_sum'].mean():.4f}")

    # ---- gradient fixture: the reference's own autograd through its training-mode `run` ----
    sd = syn.synthetic_state_dict("trained", 43)
    o, d = syn.pinhole_rays(syn.orbit_pose(30.0), 256, 256)
    sel = torch.arange(256 * 96 + 40, 256 * 160, 173)            # central rows: most rays hit the body
    ro, rd = o[sel].contiguous(), d[sel].contiguous()
    n = ro.shape[0]
    G = torch.randn(n, 3, generator=torch.Generator().manual_seed(44))     # stands in for the SDS pixel gradient
    torch.manual_seed(77)
    jitter = torch.rand(n, 64)
    net = ref_nsr.NeRFNetwork()
    net.load_state_dict(sd)
    net.train()
    torch.manual_seed(77)
    out = net.run(ro[None], rd[None], 64, 1.6, 64, None, cos_anneal_ratio=1.0, normal_epsilon_ratio=0.0,
                  render_can=True, perturb_overwrite=True)
    loss = training_loss(out, G)
    loss.backward()
    gref = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    orc = OracleNSR(sd)
    params = orc.enable_grad(sd)
    loss_o = training_loss(orc.run_grad(ro, rd, 64, 1.6, 64, jitter=jitter), G)
    loss_o.backward()
    for k, g in gref.items():
        err = float((g - params[k].grad).abs().max())
        assert err <= 2e-5 * max(1.0, float(g.abs().max())), (k, err)   # summation order of the broadcast variance
    ge = gref["encoder.embeddings"]
    nz = ge.abs().sum(1).nonzero().flatten()
    pick = nz[torch.randperm(nz.numel(), generator=torch.Generator().manual_seed(3))[:20000]].sort()[0]
    np.savez_compressed(os.path.join(GOLD, "grad_trained_jitter_64p64.npz"), rays_o=ro.numpy(), rays_d=rd.numpy(),
                        jitter=jitter.numpy(), pixel_grad=G.numpy(), loss=np.float32(float(loss)),
                        emb_rows=pick.numpy(), emb_grad=ge[pick].numpy(), emb_grad_abs_sum=np.float64(ge.abs().double().sum()),
                        emb_grad_nnz=np.int64(nz.numel()), state_checksum=syn.state_checksum(sd),
                        **{"g." + k: v.numpy() for k, v in gref.items() if k != "encoder.embeddings"})
    print(f"wrote grad_trained_jitter_64p64.npz  rays={n} loss={float(loss):.4f} nnz table rows={nz.numel()}")

    # ---- SMPL linear-blend-skinning fixture: the reference's own models/smpl.py::lbs on the synthetic model ----
    import models.smpl as ref_smpl
    md = syn.synthetic_smpl_model()
    t32 = lambda a: torch.as_tensor(np.asarray(a, np.float32))
    poses = syn.sinusoid_pose_sequence(4)
    pose, betas = torch.from_numpy(poses[2:3]), torch.zeros(1, 10)
    betas[0, 1] = 0.7
    parents = torch.as_tensor(np.asarray(md["kintree_table"][0]).astype(np.int64))
    parents[0] = -1
    posedirs = torch.zeros(207, 6890 * 3)                 # multiplied but unused by lbs (v_posed = v_shaped, smpl.py:420)
    args = (betas, pose, t32(md["v_template"])[None], t32(md["shapedirs"]), posedirs, t32(md["J_regressor"]), parents, t32(md["weights"]))
    T_ref, _, _ = ref_smpl.lbs(*args, return_T=True, concat_joints=True)
    verts_ref, J_ref = ref_smpl.lbs(*args)
    np.savez_compressed(os.path.join(GOLD, "smpl_lbs_synthetic.npz"), pose=poses[2:3], betas=betas.numpy(),
                        T_rows=np.arange(0, 6914, 37), T=T_ref[0, ::37].numpy(), posed_rows=np.arange(0, 6890, 53),
                        posed=verts_ref[0, ::53].numpy(), joints=J_ref[0].numpy())
    print("wrote smpl_lbs_synthetic.npz")

    # hash-encoder fixture through the reference's own HashEncoder.forward wrapper
    sd = syn.synthetic_state_dict("trained", 43)
    net = ref_nsr.NeRFNetwork()
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(768, 3, generator=g) * 2 - 1) * 1.6
    x[:8] = torch.tensor([[1.6, 1.6, 1.6], [-1.6, -1.6, -1.6], [0, 0, 0], [1.6, -1.6, 0.0],
                          [1.7, 0, 0], [0, -1.61, 0], [0.1, 0.2, 0.3], [1.6, 0, 1.5999999]])
    with torch.no_grad():
        feats = net.encoder(x, 1.6)
        sdf16 = net.forward_sdf(x, 1.6)
    x01 = (x + 1.6) / (2 * 1.6)
    _, ids = ohg.encode(x01, sd["encoder.embeddings"], sd["encoder.offsets"], net.encoder.per_level_scale, want_ids=True)
    np.savez_compressed(os.path.join(GOLD, "hashgrid_trained_768.npz"), x=x.numpy(), feats=feats.numpy(),
                        sdf16=sdf16.numpy(), corner_ids=ids.numpy(), state_checksum=syn.state_checksum(sd),
                        scales=ohg.level_scales(16, np.log2(net.encoder.per_level_scale), 16))
    print("wrote hashgrid_trained_768.npz")


if __name__ == "__main__":
    main()
