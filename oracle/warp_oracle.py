"""oracle/warp_oracle.py -- TEST INFRASTRUCTURE ONLY.

Restatement of the SMPL-guided inverse warp (utils/ray_utils.py:62-90, :277-294).  The reference calls
libigl (`igl.point_mesh_squared_distance`, `igl.barycentric_coordinates_tri`; igl=2.2.1, environment.yml:13),
which is NOT installed here and is not part of /root/reference: the closest-point query is therefore
restated from its geometric definition (exact brute force over all triangles, float64, Ericson's region
test) and **parity with igl's tie-breaking is unpinned** (DESIGN.md).  Everything after the query --
threshold on the SQUARED distance, barycentric blend of per-vertex 4x4s, np.linalg.inv, homogeneous apply
without renormalisation -- follows the reference lines cited below."""
import numpy as np
import torch


def closest_point_on_mesh(pts, verts, tris):
    """pts [P,3], verts [V,3], tris [F,3] -> (dist2 [P], face [P], closest [P,3], bary [P,3]) in float64.
    Ties resolve to the lowest face index (igl's choice is unknown)."""
    pts = np.asarray(pts, np.float64); verts = np.asarray(verts, np.float64)
    a, b, c = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    ab, ac = b - a, c - a
    best = np.full(pts.shape[0], np.inf); bface = np.zeros(pts.shape[0], np.int64)
    bv = np.zeros(pts.shape[0]); bw = np.zeros(pts.shape[0])
    for lo in range(0, pts.shape[0], 256):
        p = pts[lo:lo + 256, None, :]
        ap = p - a[None]
        d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
        bp = ap - ab[None]
        d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
        cp = ap - ac[None]
        d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
        vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
        with np.errstate(divide="ignore", invalid="ignore"):
            den = 1.0 / (va + vb + vc)
            v, w = vb * den, vc * den                                      # interior (default)
            m = (va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0)                # edge bc
            wbc = (d4 - d3) / ((d4 - d3) + (d5 - d6)); v = np.where(m, 1 - wbc, v); w = np.where(m, wbc, w)
            m = (vb <= 0) & (d2 >= 0) & (d6 <= 0)                          # edge ac
            v = np.where(m, 0.0, v); w = np.where(m, d2 / (d2 - d6), w)
            m = (d6 >= 0) & (d5 <= d6)                                     # vertex c
            v = np.where(m, 0.0, v); w = np.where(m, 1.0, w)
            m = (vc <= 0) & (d1 >= 0) & (d3 <= 0)                          # edge ab
            v = np.where(m, d1 / (d1 - d3), v); w = np.where(m, 0.0, w)
            m = (d3 >= 0) & (d4 <= d3)                                     # vertex b
            v = np.where(m, 1.0, v); w = np.where(m, 0.0, w)
            m = (d1 <= 0) & (d2 <= 0)                                      # vertex a  (highest priority, tested first)
            v = np.where(m, 0.0, v); w = np.where(m, 0.0, w)
        q = ap - (ab[None] * v[..., None] + ac[None] * w[..., None])
        dist2 = (q * q).sum(-1)
        j = dist2.argmin(1)
        r = np.arange(j.shape[0])
        best[lo:lo + 256], bface[lo:lo + 256], bv[lo:lo + 256], bw[lo:lo + 256] = dist2[r, j], j, v[r, j], w[r, j]
    closest = a[bface] + ab[bface] * bv[:, None] + ac[bface] * bw[:, None]
    return best, bface, closest, np.stack([1 - bv - bw, bv, bw], -1)


def warp_samples_to_canonical(pts, verts, faces, T, threshold=0.05):
    """utils/ray_utils.py:62-90.  pts [R,S,3] -> (can_pts [R,S,3] f64, mask [R,S] bool, closest, face, dist2)."""
    R, S, _ = pts.shape
    flat = np.asarray(pts, np.float64).reshape(-1, 3)
    tris = np.asarray(faces)[:, :3]
    dist2, face, closest, bary = closest_point_on_mesh(flat, verts, tris)
    mask = dist2 < threshold                                              # threshold on the SQUARED distance (:74)
    T_interp = (np.asarray(T, np.float64)[tris[face]] * bary[..., None, None]).sum(1)     # (:80)
    hom = np.concatenate([flat, np.ones_like(flat[:, :1])], -1)
    can = (np.linalg.inv(T_interp) @ hom[..., None])[:, :3, 0]            # (:81-84) no division by w
    return can.reshape(R, S, 3), mask.reshape(R, S), closest.reshape(R, S, 3), face.reshape(R, S), dist2.reshape(R, S)


def geometry_guided_near_far(orig, dirs, vert, geo_threshold=0.05):
    """utils/ray_utils.py:277-294 (torch variant): near/far from the radius-`geo_threshold` spheres around
    the posed vertices; +-inf where the ray pierces none."""
    vert = torch.as_tensor(vert, dtype=torch.float32)
    ov = vert[None] - orig[:, None]
    z0 = (ov * dirs[:, None]).sum(-1)
    dz = torch.sqrt(geo_threshold ** 2 - (torch.norm(ov, dim=2) ** 2 - z0 ** 2))
    near = z0 - dz
    near[near != near] = float("inf")
    far = z0 + dz
    far[far != far] = float("-inf")
    return near.min(dim=1)[0], far.max(dim=1)[0]
