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


def _ericson(ap, ab, ac):
    """Closest point of triangles (a, a+ab, a+ac) to points given as ap = p - a (broadcast shapes [..., 3]): (dist2, v, w)."""
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = ap - ab
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = ap - ac
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
    q = ap - (ab * v[..., None] + ac * w[..., None])
    return (q * q).sum(-1), v, w


def closest_point_on_mesh_pruned(pts, verts, tris, k=64):
    """Same result as closest_point_on_mesh (exact, lowest face index among ties), for large point sets: the candidates of a
    point are the k triangles with the nearest centroids (scipy cKDTree); a point is ACCEPTED only if its best distance is
    provably not beaten by any other triangle (every non-candidate's centroid is at least d_k away, so its surface is at least
    d_k - r_max away, r_max = the largest centroid-to-vertex distance); the rest go through the brute-force scan."""
    from scipy.spatial import cKDTree
    pts = np.asarray(pts, np.float64); verts = np.asarray(verts, np.float64)
    a, b, c = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    ab, ac = b - a, c - a
    cen = (a + b + c) / 3.0
    r_max = float(np.sqrt(np.maximum(((a - cen) ** 2).sum(-1), np.maximum(((b - cen) ** 2).sum(-1), ((c - cen) ** 2).sum(-1)))).max())
    P = pts.shape[0]
    best = np.empty(P); bface = np.empty(P, np.int64); bv = np.empty(P); bw = np.empty(P)
    tree = cKDTree(cen)
    todo = np.arange(P)
    for kk in (k, 8 * k):                                            # far points need a wider candidate set before they verify
        kk = min(kk, tris.shape[0])
        left = []
        chunk = max(256, (1 << 19) // kk)
        for lo in range(0, todo.size, chunk):
            sel = todo[lo:lo + chunk]
            p = pts[sel]
            dk, idx = tree.query(p, k=kk)
            idx = np.sort(idx, axis=1)                               # ascending face ids: argmin then picks the lowest id among ties
            d2, v, w = _ericson(p[:, None, :] - a[idx], ab[idx], ac[idx])
            j = d2.argmin(1)
            r = np.arange(j.shape[0])
            best[sel], bface[sel], bv[sel], bw[sel] = d2[r, j], idx[r, j], v[r, j], w[r, j]
            unsure = np.sqrt(d2[r, j]) > dk[:, -1] - r_max           # a triangle outside the candidate set could be as close
            left.append(sel[unsure])
        todo = np.concatenate(left) if left else np.zeros(0, np.int64)
        if todo.size == 0:
            break
    if todo.size:
        d2, f, _, bary = closest_point_on_mesh(pts[todo], verts, tris)
        best[todo], bface[todo], bv[todo], bw[todo] = d2, f, bary[:, 1], bary[:, 2]
    closest = a[bface] + ab[bface] * bv[:, None] + ac[bface] * bw[:, None]
    return best, bface, closest, np.stack([1 - bv - bw, bv, bw], -1)


def closest_point_on_mesh_torch(pts, verts, tris, device):
    """closest_point_on_mesh as a float64 torch brute-force scan on `device` -- the same region test over ALL triangles, the
    same tie rule (first minimum = lowest face index).  Lets the -m gpu tests run the oracle on thousands of rays: the checker
    may use the GPU's fp64 units, it is still an exhaustive restatement and never part of the product."""
    dev = torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    P = torch.as_tensor(np.asarray(pts, np.float64), **f64)
    V = torch.as_tensor(np.asarray(verts, np.float64), **f64)
    F = torch.as_tensor(np.asarray(tris), device=dev).long()
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    ab, ac = (b - a)[None], (c - a)[None]
    outs = []
    for lo in range(0, P.shape[0], 512):
        ap = P[lo:lo + 512, None, :] - a[None]
        d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
        bp = ap - ab
        d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
        cp = ap - ac
        d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
        vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
        den = 1.0 / (va + vb + vc)
        v, w = vb * den, vc * den
        zero, one = torch.zeros_like(v), torch.ones_like(v)
        m = (va <= 0) & (d4 - d3 >= 0) & (d5 - d6 >= 0)
        wbc = (d4 - d3) / ((d4 - d3) + (d5 - d6)); v = torch.where(m, 1 - wbc, v); w = torch.where(m, wbc, w)
        m = (vb <= 0) & (d2 >= 0) & (d6 <= 0)
        v = torch.where(m, zero, v); w = torch.where(m, d2 / (d2 - d6), w)
        m = (d6 >= 0) & (d5 <= d6)
        v = torch.where(m, zero, v); w = torch.where(m, one, w)
        m = (vc <= 0) & (d1 >= 0) & (d3 <= 0)
        v = torch.where(m, d1 / (d1 - d3), v); w = torch.where(m, zero, w)
        m = (d3 >= 0) & (d4 <= d3)
        v = torch.where(m, one, v); w = torch.where(m, zero, w)
        m = (d1 <= 0) & (d2 <= 0)
        v = torch.where(m, zero, v); w = torch.where(m, zero, w)
        q = ap - (ab * v[..., None] + ac * w[..., None])
        dist2 = (q * q).sum(-1)
        j = dist2.argmin(1)
        r = torch.arange(j.shape[0], device=dev)
        outs.append((dist2[r, j], j, v[r, j], w[r, j]))
    best, bface, bv, bw = (torch.cat([o[i] for o in outs]).cpu().numpy() for i in range(4))
    a, ab, ac = a.cpu().numpy(), ab[0].cpu().numpy(), ac[0].cpu().numpy()
    closest = a[bface] + ab[bface] * bv[:, None] + ac[bface] * bw[:, None]
    return best, bface, closest, np.stack([1 - bv - bw, bv, bw], -1)


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


def warp_samples_to_canonical(pts, verts, faces, T, threshold=0.05, device=None):
    """utils/ray_utils.py:62-90.  pts [R,S,3] -> (can_pts [R,S,3] f64, mask [R,S] bool, closest, face, dist2).
    `device` ("cuda"): run the exhaustive closest-point scan in float64 torch there (large point sets of the GPU tests)."""
    R, S, _ = pts.shape
    flat = np.asarray(pts, np.float64).reshape(-1, 3)
    tris = np.asarray(faces)[:, :3]
    if device is not None:
        dist2, face, closest, bary = closest_point_on_mesh_torch(flat, verts, tris, device)
    else:
        query = closest_point_on_mesh_pruned if flat.shape[0] > 20000 else closest_point_on_mesh
        dist2, face, closest, bary = query(flat, verts, tris)
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
