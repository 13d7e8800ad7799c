"""Device-side counterparts of the reference's utils/ray_utils.py warp helpers.

The reference moves every sample point to the host, queries libigl's AABB tree, inverts a blended 4x4 per
point in numpy and moves the result back (utils/ray_utils.py:62-90; models/instant_nsr.py:166-172,198-203).
These functions keep the same call shape but take and return CUDA tensors and run as single kernels."""
import ctypes

import numpy as np
import torch

from .. import _lib
from .constant import DEFAULT_GEO_THRESH


SORT_QUERIES_FROM = 1 << 14     # below this the sort costs more than the divergence it removes
RAY_COHERENT = False            # opt-in: walk each ray's samples in order, seeding every search with the previous result.  Measured
                                # slower than Morton-sorted per-point queries (7.3 vs 4.2 ms per 2-4 M points: fewer, longer threads)


class PosedMesh:
    """Per-frame device copy of the posed SMPL surface: vertices, faces, per-vertex transforms and the
    64-byte triangle records (vertex, two edges, bounding sphere, ids) the warp kernel walks."""

    def __init__(self, verts, faces, Ts, device):
        self.verts = torch.as_tensor(np.asarray(verts), dtype=torch.float32).to(device).contiguous()
        f = torch.as_tensor(np.asarray(faces)).to(torch.int32).to(device)
        # Sort the triangles along a Morton curve of their centroids so that every 64 consecutive records (one
        # cluster of the kernel's branch-and-bound search) are spatially compact.  `order` maps sorted -> original ids.
        cen = self.verts[f[:, :3].long()].mean(1)
        q = ((cen - cen.min(0)[0]) / (cen.max(0)[0] - cen.min(0)[0] + 1e-12) * 1023.0).long().clamp_(0, 1023)

        def spread(v):                       # 10 bits -> every third bit
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            return (v | (v << 2)) & 0x09249249
        self.order = torch.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2))
        f = f[self.order].contiguous()
        self.faces = f
        T = torch.as_tensor(np.asarray(Ts), dtype=torch.float32)
        if float(T[:, 3, :3].abs().max()) != 0.0:
            raise RuntimeError("per-vertex transforms must be affine up to a homogeneous scale: last rows (0,0,0,c)")
        self.Ts = T.to(device).contiguous()
        self.n_faces = int(f.shape[0])
        L = _lib.lib()
        self.records = torch.empty(int(L.ac_warp_mesh_bytes(self.n_faces)), dtype=torch.uint8, device=device)
        _lib.check(L.ac_warp_prepare_mesh(_lib.ptr(self.verts), _lib.ptr(f), int(f.shape[1]), self.n_faces,
                                          _lib.ptr(self.records), _lib.stream_ptr()), "ac_warp_prepare_mesh")


def warp_samples_to_canonical(pts, verts, faces, T, threshold=0.2, mesh: PosedMesh = None, return_query=False, product=False,
                              masked_only=False):
    """pts [num_rays, num_samples, 3] (CUDA) -> can_pts, can_dirs, closest, mask -- the reference's return tuple
    (utils/ray_utils.py:62-90).  `can_dirs` is computed like the reference does although nothing downstream
    reads it (models/instant_nsr.py:203,208).  `product=True` (the render path): only what `run` consumes -- (can_pts
    [R,S,3], mask [R,S] as the kernel's 0/1 floats) -- and no torch arithmetic.  `masked_only` (with `product`): the caller
    multiplies alpha by the mask, so samples farther than sqrt(threshold) from the surface need no closest point -- the search
    is bounded by the threshold, those samples return mask 0 and can_pts = pts; everything inside is unchanged."""
    assert pts.dim() == 3 and pts.shape[-1] == 3, 'pts should have shape [num_rays, num_samples, 3]'
    if mesh is None:
        mesh = PosedMesh(verts, faces, T, pts.device)
    R, S, _ = pts.shape
    flat = pts.reshape(-1, 3).float().contiguous()
    n = flat.shape[0]
    can = torch.empty_like(flat); mask = torch.empty(n, device=flat.device)
    closest = torch.empty_like(flat)
    face = torch.empty(n, dtype=torch.int32, device=flat.device) if return_query else None
    dist2 = torch.empty(n, device=flat.device) if return_query else None
    if RAY_COHERENT and S >= 8:
        # consecutive samples of a ray are neighbours: each thread walks 8 of them, seeding every search with the previous
        # result (ac_warp_samples_to_canonical_rays) -- bit-identical to the per-point search, no sort needed (opt-in, see above)
        _lib.check(_lib.lib().ac_warp_samples_to_canonical_rays(_lib.ptr(flat), R, S, _lib.ptr(mesh.records), mesh.n_faces, _lib.ptr(mesh.Ts),
                                                                float(threshold), _lib.ptr(can), _lib.ptr(mask), _lib.ptr(closest),
                                                                _lib.ptr(face), _lib.ptr(dist2), _lib.stream_ptr()),
                   "ac_warp_samples_to_canonical_rays")
    else:
        order = None
        if n >= SORT_QUERIES_FROM:
            # Visit the queries along a Morton curve: a warp's 32 points then share boxes and triangles (11.7 -> ~30 active
            # lanes per instruction in the search); results land at the original indices.
            keys = torch.empty(n, dtype=torch.int32, device=flat.device)
            _lib.check(_lib.lib().ac_warp_query_keys(_lib.ptr(flat), n, _lib.ptr(mesh.records), mesh.n_faces, 0.25, _lib.ptr(keys),
                                                     _lib.stream_ptr()), "ac_warp_query_keys")
            order = torch.sort(keys)[1].to(torch.int32)
        if product and masked_only:
            _lib.check(_lib.lib().ac_warp_samples_to_canonical_masked(_lib.ptr(flat), _lib.ptr(order), n, _lib.ptr(mesh.records), mesh.n_faces,
                                                                      _lib.ptr(mesh.Ts), float(threshold), _lib.ptr(can), _lib.ptr(mask),
                                                                      _lib.stream_ptr()), "ac_warp_samples_to_canonical_masked")
            return can.reshape(R, S, 3), mask.reshape(R, S)
        _lib.check(_lib.lib().ac_warp_samples_to_canonical_ordered(_lib.ptr(flat), _lib.ptr(order), n, _lib.ptr(mesh.records), mesh.n_faces,
                                                                   _lib.ptr(mesh.Ts), float(threshold), _lib.ptr(can), _lib.ptr(mask),
                                                                   _lib.ptr(closest), _lib.ptr(face), _lib.ptr(dist2), _lib.stream_ptr()),
                   "ac_warp_samples_to_canonical")
    can = can.reshape(R, S, 3)
    if product:
        return can, mask.reshape(R, S)
    dirs = can[:, 1:] - can[:, :-1]
    dirs = torch.cat([dirs, dirs[:, -1:]], dim=1)
    dirs = dirs / torch.linalg.norm(dirs, dim=2, keepdim=True)
    out = (can, dirs, closest.reshape(R, S, 3), mask.reshape(R, S) > 0.5)
    return out + (mesh.order[face.long()].to(torch.int32).reshape(R, S), dist2.reshape(R, S)) if return_query else out


def geometry_guided_near_far(orig, dir, vert, geo_threshold=DEFAULT_GEO_THRESH, bound=None):
    """near/far [n] from the radius-`geo_threshold` spheres around the posed vertices
    (utils/ray_utils.py:277-294).  With `bound` given, rays that pierce no sphere fall back to the cube
    intersection as models/instant_nsr.py:147-153 does; without it they get +-inf like the reference helper."""
    vert = torch.as_tensor(np.asarray(vert) if not isinstance(vert, torch.Tensor) else vert, dtype=torch.float32).to(orig.device).contiguous()
    o, d = orig.reshape(-1, 3).float().contiguous(), dir.reshape(-1, 3).float().contiguous()
    nf = torch.empty(o.shape[0], 2, device=o.device)
    _lib.check(_lib.lib().ac_mesh_guided_near_far(_lib.ptr(o), _lib.ptr(d), o.shape[0], _lib.ptr(vert), vert.shape[0],
                                                  float(geo_threshold), float(bound if bound is not None else 1e30), _lib.ptr(nf),
                                                  _lib.stream_ptr()), "ac_mesh_guided_near_far")
    if bound is None:
        raise NotImplementedError("pass bound: the cube fallback is fused into the kernel")
    return nf[:, 0], nf[:, 1]
