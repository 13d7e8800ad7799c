"""Synthetic stand-ins for the assets the reference needs but does not ship
(checkpoints are Google-Drive links, readme.md:55,72): deterministic checkpoints in the
reference's state-dict layout and the canonical-orbit camera rays of
render_canonical.py.  Host-side numpy/torch only."""
import math

import numpy as np
import torch

STATE_KEYS = (
    "encoder.embeddings", "encoder.offsets",
    "sdf_net.0.bias", "sdf_net.0.weight_g", "sdf_net.0.weight_v",
    "sdf_net.1.bias", "sdf_net.1.weight_g", "sdf_net.1.weight_v",
    "color_net.0.weight_g", "color_net.0.weight_v",
    "color_net.1.weight_g", "color_net.1.weight_v",
    "color_net.2.weight_g", "color_net.2.weight_v",
    "deviation_net.variance",
)


def hash_offsets(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048):
    """Offset table + per-level scale with the sizing rule of encoder/hashencoder/hashgrid.py:84-108."""
    pls = float(np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1)))
    cap, offs, total = 2 ** log2_hashmap_size, [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * pls ** i))
        offs.append(total)
        total += min(cap, (res + 1) ** input_dim)
    offs.append(total)
    return torch.tensor(offs, dtype=torch.int32), pls


def synthetic_state_dict(kind="trained", seed=43):
    """kind='init': statistics of the reference's fresh initialisation (geometric init,
    models/instant_nsr.py:537-553; table U(-1e-4,1e-4), hashgrid.py:119-121; variance 0.3).
    kind='trained': a "trained-like" field -- a bumpy sphere of radius ~0.5 whose hash features
    matter, with a sharp NeuS transition (variance 0.6 -> inv_s ~ 403)."""
    g = torch.Generator().manual_seed(seed)
    offsets, _ = hash_offsets()
    n = int(offsets[-1])
    sd = {"encoder.offsets": offsets}
    u = torch.rand(n, 2, generator=g) * 2 - 1
    if kind == "init":
        sd["encoder.embeddings"] = u * 1e-4
    else:
        amp = torch.empty(n)
        for l in range(16):
            amp[int(offsets[l]):int(offsets[l + 1])] = 0.1 * 0.85 ** l
        sd["encoder.embeddings"] = u * amp[:, None]
    v0 = torch.zeros(64, 35)
    v0[:, :3] = torch.randn(64, 3, generator=g) * (math.sqrt(2) / math.sqrt(64))
    if kind != "init":
        v0[:, 3:] = torch.randn(64, 32, generator=g) * 0.1
    v1 = torch.randn(16, 64, generator=g) * (1e-4 if kind == "init" else 0.05)
    v1[0] = math.sqrt(math.pi) / math.sqrt(64) + torch.randn(64, generator=g) * 1e-4
    if kind == "init":
        v1 = math.sqrt(math.pi) / math.sqrt(64) + torch.randn(16, 64, generator=g) * 1e-4
    b1 = torch.zeros(16)
    if kind != "init":
        b1[0] = -0.5
        b1[1:] = torch.randn(15, generator=g) * 0.1
    sd["sdf_net.0.weight_v"], sd["sdf_net.0.bias"] = v0, torch.zeros(64) if kind == "init" else torch.randn(64, generator=g) * 0.01
    sd["sdf_net.1.weight_v"], sd["sdf_net.1.bias"] = v1, b1
    for i, (o, k) in enumerate(((64, 21), (64, 64), (3, 64))):
        bnd = 1.0 / math.sqrt(k)                       # nn.Linear default init range
        sd[f"color_net.{i}.weight_v"] = (torch.rand(o, k, generator=g) * 2 - 1) * bnd * (1.0 if kind == "init" else 2.0)
    for name in ("sdf_net.0", "sdf_net.1", "color_net.0", "color_net.1", "color_net.2"):
        v = sd[name + ".weight_v"]
        gain = v.norm(dim=1, keepdim=True)             # weight_norm initialises g = ||v||
        if kind != "init":
            gain = gain * (1.0 + 0.1 * (torch.rand(v.shape[0], 1, generator=g) - 0.5))
        sd[name + ".weight_g"] = gain
    sd["deviation_net.variance"] = torch.tensor(0.3 if kind == "init" else 0.6)
    return {k: sd[k] for k in STATE_KEYS}


def state_checksum(sd):
    """Order-independent fingerprint so golden fixtures can detect RNG drift."""
    tot = 0.0
    for k in STATE_KEYS:
        t = sd[k].double().flatten()
        tot += float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).remainder(97.0)).sum())
    return tot


def orbit_pose(angle_deg, dist=1.7, center=(0.0, 0.0, 0.0)):
    """Camera-to-world of a y-up orbit camera looking at `center` (the geometry of
    utils/render_utils.py:137-154 default_360_path; +z is the front view at angle 0)."""
    a = math.radians(angle_deg)
    eye = np.array([dist * math.sin(a), 0.0, dist * math.cos(a)]) + np.asarray(center)
    fwd = np.asarray(center) - eye
    fwd = fwd / np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, -up, fwd, eye   # x right, y down, z forward
    return c2w


def pinhole_rays(c2w, width, height):
    """Row-major pixel-centre rays of a pinhole with f = 0.78125*W, c = W/2
    (render_canonical.py:60-71; utils/ray_utils.py:25-37): float32 origins and unit dirs."""
    fx, fy, cx, cy = 0.78125 * width, 0.78125 * height, width / 2, height / 2
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    cam = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs, dtype=np.float64)], -1).reshape(-1, 3)
    world = (cam @ c2w[:3, :3].T + c2w[:3, 3]).astype(np.float32)
    origin = np.broadcast_to(c2w[:3, 3].astype(np.float32), world.shape).copy()
    d = world - origin
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    return torch.from_numpy(origin), torch.from_numpy(d.astype(np.float32))


def synthetic_body(pose_seed=46, amplitude=0.35):
    """SMPL-shaped stand-in (the licensed SMPL model and smpl_uv.obj are not in the repo, readme.md:41-59):
    a closed genus-0 mesh with SMPL's exact counts -- 6890 vertices, 13776 faces (82 rings x 84 segments + 2
    poles) -- a 24-joint chain with smooth skinning weights, and per-vertex rest->pose 4x4 transforms built the
    way render_warp.py:127-222 composes them (blend of rigid joint transforms, times diag(1/0.9) INCLUDING the
    homogeneous entry).  Returns dict(rest_verts [6890,3], world_verts [6890,3] f32, faces [13776,6] int32 (first 3
    columns = vertex ids, like utils.read_obj), Ts [6914,4,4] f32)."""
    R, S = 82, 84
    rng = np.random.default_rng(pose_seed)
    ys = np.cos(np.linspace(0, np.pi, R + 2))[1:-1]                 # ring heights in (-1, 1), top to bottom
    ang = np.linspace(0, 2 * np.pi, S, endpoint=False)
    prof = np.sqrt(np.clip(1 - ys ** 2, 0, None)) * (1.0 + 0.25 * np.cos(3.0 * np.pi * ys))   # waist / shoulders
    rings = np.stack([np.outer(prof, np.cos(ang)) * 0.28, np.repeat(ys[:, None], S, 1) * 0.85, np.outer(prof, np.sin(ang)) * 0.17], -1)
    verts = np.concatenate([[[0, 0.85, 0]], rings.reshape(-1, 3), [[0, -0.85, 0]]]).astype(np.float64)
    f = []
    top, bot = 0, 1 + R * S
    idx = lambda r, s: 1 + r * S + (s % S)
    for s in range(S):
        f.append((top, idx(0, s + 1), idx(0, s)))
        f.append((bot, idx(R - 1, s), idx(R - 1, s + 1)))
        for r in range(R - 1):
            f.append((idx(r, s), idx(r, s + 1), idx(r + 1, s)))
            f.append((idx(r, s + 1), idx(r + 1, s + 1), idx(r + 1, s)))
    faces = np.array(f, dtype=np.int32)
    assert verts.shape[0] == 6890 and faces.shape[0] == 13776
    # 24-joint chain along y, smooth weights (4 neighbouring joints per vertex)
    joints = np.stack([np.zeros(24), np.linspace(0.8, -0.8, 24), np.zeros(24)], -1)
    dj = np.abs(verts[:, 1:2] - joints[None, :, 1])
    w = np.exp(-(dj / 0.08) ** 2)
    w[np.arange(6890)[:, None], np.argsort(-w, 1)[:, 4:]] = 0.0
    w /= w.sum(1, keepdims=True)
    # forward kinematics: small random axis-angle per joint, chained
    A = np.zeros((24, 4, 4)); G = np.eye(4)
    for j in range(24):
        aa = rng.normal(size=3) * amplitude * (0.3 if j else 0.1)
        th = np.linalg.norm(aa) + 1e-12
        k = aa / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        Rm = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        L = np.eye(4); L[:3, :3] = Rm; L[:3, 3] = joints[j] - Rm @ joints[j]      # rotate about the joint
        G = G @ L
        A[j] = G
    T_v = np.einsum("vj,jab->vab", w, A)                                  # per-vertex blended transform
    T_all = np.concatenate([T_v, A], 0)                                    # 6890 vertices + 24 joints ("concat_joints")
    world = (T_v[:, :3, :3] @ verts[:, :, None])[:, :, 0] + T_v[:, :3, 3]
    Ts = T_all @ (np.eye(4) / 0.9)
    faces6 = np.concatenate([faces, faces], 1)
    return dict(rest_verts=verts.astype(np.float32), world_verts=world.astype(np.float32), faces=faces6,
                Ts=Ts.astype(np.float32))


def synthetic_smpl_model(seed=45):
    """SMPL-shaped model dict (fields of SMPL_NEUTRAL.pkl used by models/smpl.py) on the synthetic body:
    6890-vertex template, 10 smooth shape directions, a 24-joint regressor (mean of the ring nearest each
    joint height), SMPL's kinematic tree, sparse skinning weights (4 non-zeros per vertex)."""
    body = synthetic_body()
    V = body["rest_verts"].astype(np.float64)
    rng = np.random.default_rng(seed)
    joints_y = np.linspace(0.8, -0.8, 24)
    dj = np.abs(V[:, 1:2] - joints_y[None])
    w = np.exp(-(dj / 0.08) ** 2)
    w[np.arange(6890)[:, None], np.argsort(-w, 1)[:, 4:]] = 0.0
    w /= w.sum(1, keepdims=True)
    Jr = np.exp(-(dj.T / 0.03) ** 2)
    Jr /= Jr.sum(1, keepdims=True)
    dirs = np.stack([np.sin((k + 1) * V[:, 1:2] * 2.0 + rng.uniform(0, 6.28)) * V * 0.03 for k in range(10)], -1)   # [V,3,10]
    return dict(v_template=V, shapedirs=dirs, J_regressor=Jr, weights=w, f=body["faces"][:, :3],
                kintree_table=np.stack([np.array([2 ** 32 - 1] + [p for p in (0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)]),
                                        np.arange(24)]))


def sinusoid_pose_sequence(n_frames=120, amplitude=0.5, seed=46):
    """[F,72] smooth axis-angle sequence standing in for an AMASS-SFU clip (convert_amass.py:5-17 layout)."""
    rng = np.random.default_rng(seed)
    phase, freq = rng.uniform(0, 6.28, (24, 3)), rng.uniform(0.5, 2.0, (24, 3))
    t = np.linspace(0, 2 * np.pi, n_frames)[:, None, None]
    p = amplitude * 0.3 * np.sin(freq[None] * t + phase[None])
    p[:, 0] *= 0.2
    return p.reshape(n_frames, 72).astype(np.float32)
