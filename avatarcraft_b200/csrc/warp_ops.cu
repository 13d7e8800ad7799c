// warp_ops.cu -- the SMPL-guided inverse warp of sample points, on the GPU.
//
// Reference: utils/ray_utils.py:62-90 warp_samples_to_canonical -- sample points go device->host,
// libigl's AABB-tree closest-point query runs on the CPU, numpy inverts a blended 4x4 per point, and the
// result goes back to the device, twice per ray batch (models/instant_nsr.py:166-172,198-203).
// Here: one kernel per point set, no host round trip.  Closest point on the posed mesh is an EXACT
// branch-and-bound search (Ericson's region test per triangle): triangles arrive sorted along a Morton curve
// (host side, utils/ray_utils.PosedMesh); 16 consecutive records form a leaf box, 8 leaves a 128-triangle box,
// 8 of those a 1024-triangle box (an implicit 3-level AABB tree rebuilt per pose in 3 tiny launches).  A query
// descends greedily to the nearest leaf for a first bound, then visits only the boxes (and, inside a leaf, only
// the triangles, by their bounding spheres) that can still beat or tie the current best.
// utils/ray_utils.py:277-294 geometry_guided_near_far becomes a warp-per-ray reduction over vertices.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"

namespace {

struct __align__(16) TriRec {       // 64 bytes = 4 x float4; the bounding sphere comes first so a rejected triangle costs one load
    float cx, cy, cz, r;            // bounding sphere
    float ax, ay, az; int32_t i0;   // vertex a, vertex ids
    float abx, aby, abz; int32_t i1;// edge b-a
    float acx, acy, acz; int32_t i2;// edge c-a
};

__global__ void __launch_bounds__(256) mesh_prepare_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces,
                                                           uint32_t face_stride, uint32_t n_faces, TriRec* __restrict__ out) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const int32_t i0 = faces[(size_t)f * face_stride], i1 = faces[(size_t)f * face_stride + 1], i2 = faces[(size_t)f * face_stride + 2];
    const float ax = verts[3 * i0], ay = verts[3 * i0 + 1], az = verts[3 * i0 + 2];
    const float bx = verts[3 * i1], by = verts[3 * i1 + 1], bz = verts[3 * i1 + 2];
    const float cx = verts[3 * i2], cy = verts[3 * i2 + 1], cz = verts[3 * i2 + 2];
    TriRec t;
    t.ax = ax; t.ay = ay; t.az = az;
    t.abx = bx - ax; t.aby = by - ay; t.abz = bz - az;
    t.acx = cx - ax; t.acy = cy - ay; t.acz = cz - az;
    const float mx = (ax + bx + cx) / 3.0f, my = (ay + by + cy) / 3.0f, mz = (az + bz + cz) / 3.0f;
    const float ra = (ax - mx) * (ax - mx) + (ay - my) * (ay - my) + (az - mz) * (az - mz);
    const float rb = (bx - mx) * (bx - mx) + (by - my) * (by - my) + (bz - mz) * (bz - mz);
    const float rc = (cx - mx) * (cx - mx) + (cy - my) * (cy - my) + (cz - mz) * (cz - mz);
    t.cx = mx; t.cy = my; t.cz = mz;
    t.r = sqrtf(fmaxf(ra, fmaxf(rb, rc))) * 1.0001f + 1e-7f;   // conservative
    t.i0 = i0; t.i1 = i1; t.i2 = i2;
    out[f] = t;
}

// Closest point on triangle (a, a+ab, a+ac) to p: returns squared distance and barycentrics (b0,b1,b2).
__device__ __forceinline__ float closest_on_triangle(const TriRec& t, float px, float py, float pz, float& b1, float& b2) {
    const float apx = px - t.ax, apy = py - t.ay, apz = pz - t.az;
    const float d1 = t.abx * apx + t.aby * apy + t.abz * apz;
    const float d2 = t.acx * apx + t.acy * apy + t.acz * apz;
    float v, w;
    do {
        if (d1 <= 0.f && d2 <= 0.f) { v = 0.f; w = 0.f; break; }                          // vertex a
        const float bpx = apx - t.abx, bpy = apy - t.aby, bpz = apz - t.abz;
        const float d3 = t.abx * bpx + t.aby * bpy + t.abz * bpz;
        const float d4 = t.acx * bpx + t.acy * bpy + t.acz * bpz;
        if (d3 >= 0.f && d4 <= d3) { v = 1.f; w = 0.f; break; }                            // vertex b
        const float vc = d1 * d4 - d3 * d2;
        if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { v = d1 / (d1 - d3); w = 0.f; break; }   // edge ab
        const float cpx = apx - t.acx, cpy = apy - t.acy, cpz = apz - t.acz;
        const float d5 = t.abx * cpx + t.aby * cpy + t.abz * cpz;
        const float d6 = t.acx * cpx + t.acy * cpy + t.acz * cpz;
        if (d6 >= 0.f && d5 <= d6) { v = 0.f; w = 1.f; break; }                            // vertex c
        const float vb = d5 * d2 - d1 * d6;
        if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { v = 0.f; w = d2 / (d2 - d6); break; }   // edge ac
        const float va = d3 * d6 - d5 * d4;
        if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {                            // edge bc
            w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); v = 1.f - w; break;
        }
        const float den = 1.0f / (va + vb + vc);                                            // interior
        v = vb * den; w = vc * den;
    } while (false);
    const float qx = apx - (t.abx * v + t.acx * w), qy = apy - (t.aby * v + t.acy * w), qz = apz - (t.abz * v + t.acz * w);
    b1 = v; b2 = w;
    return qx * qx + qy * qy + qz * qz;
}

constexpr uint32_t kCluster = 16;     // triangles per leaf box (64: 69 ms per 256x256 animate frame with a flat list)
constexpr uint32_t kFan = 8;          // children per inner box: leaves -> 128-triangle boxes -> 1024-triangle boxes

struct __align__(16) Box { float lx, ly, lz, pad0, hx, hy, hz, pad1; };   // 32 bytes = 2 x float4

struct MeshView {                     // layout of the caller's `mesh` buffer (ac_warp_mesh_bytes)
    const TriRec* tris; const Box* l0; const Box* l1; const Box* l2;
    uint32_t n_faces, n0, n1, n2;
};
__host__ __device__ inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
__host__ inline MeshView mesh_view(const void* mesh, uint32_t n_faces) {
    MeshView m;
    m.n_faces = n_faces; m.n0 = ceil_div(n_faces, kCluster); m.n1 = ceil_div(m.n0, kFan); m.n2 = ceil_div(m.n1, kFan);
    m.tris = reinterpret_cast<const TriRec*>(mesh);
    m.l0 = reinterpret_cast<const Box*>(m.tris + n_faces); m.l1 = m.l0 + m.n0; m.l2 = m.l1 + m.n1;
    return m;
}

// Leaf boxes: one thread per 16 consecutive (Morton-ordered) triangles.  Padded so that the triangle as the
// query evaluates it (a, a+ab, a+ac with rounded edges) is inside.
__global__ void __launch_bounds__(128) mesh_leaf_box_kernel(const TriRec* __restrict__ tris, uint32_t n_faces, Box* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = c * kCluster;
    if (lo >= n_faces) return;
    const uint32_t hi = min(n_faces, lo + kCluster);
    float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t k = lo; k < hi; ++k) {
        const TriRec t = tris[k];
        const float v[3][3] = {{t.ax, t.ay, t.az}, {t.ax + t.abx, t.ay + t.aby, t.az + t.abz}, {t.ax + t.acx, t.ay + t.acy, t.az + t.acz}};
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int d = 0; d < 3; ++d) { l[d] = fminf(l[d], v[q][d]); h[d] = fmaxf(h[d], v[q][d]); }
    }
    Box b;
    const float pad = 1e-6f + 1e-5f * fmaxf(h[0] - l[0], fmaxf(h[1] - l[1], h[2] - l[2]));
    b.lx = l[0] - pad; b.ly = l[1] - pad; b.lz = l[2] - pad; b.hx = h[0] + pad; b.hy = h[1] + pad; b.hz = h[2] + pad;
    b.pad0 = b.pad1 = 0.f;
    out[c] = b;
}
// Inner boxes: union of kFan children.
__global__ void __launch_bounds__(128) mesh_inner_box_kernel(const Box* __restrict__ child, uint32_t n_child, Box* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = c * kFan;
    if (lo >= n_child) return;
    const uint32_t hi = min(n_child, lo + kFan);
    Box b = child[lo];
    for (uint32_t k = lo + 1; k < hi; ++k) {
        const Box o = child[k];
        b.lx = fminf(b.lx, o.lx); b.ly = fminf(b.ly, o.ly); b.lz = fminf(b.lz, o.lz);
        b.hx = fmaxf(b.hx, o.hx); b.hy = fmaxf(b.hy, o.hy); b.hz = fmaxf(b.hz, o.hz);
    }
    out[c] = b;
}

// Squared distance from p to a box (0 inside): a lower bound of the squared distance to anything in it.
__device__ __forceinline__ float box_dist2(const Box* __restrict__ b, float px, float py, float pz) {
    const float4 lo = __ldg(reinterpret_cast<const float4*>(b)), hi = __ldg(reinterpret_cast<const float4*>(b) + 1);
    const float dx = fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f),
                dz = fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// 30-bit Morton key of a query point inside the mesh's bounding box grown by `margin` (clamped outside): sorting
// the queries by it puts 32 spatially adjacent points in a warp.
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v = (v | (v << 16)) & 0x030000FFu; v = (v | (v << 8)) & 0x0300F00Fu; v = (v | (v << 4)) & 0x030C30C3u;
    return (v | (v << 2)) & 0x09249249u;
}
__global__ void __launch_bounds__(256) query_keys_kernel(const float* __restrict__ pts, uint32_t n, const MeshView m, float margin,
                                                         int32_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t s = 0; s < m.n2; ++s) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(m.l2 + s)), hi = __ldg(reinterpret_cast<const float4*>(m.l2 + s) + 1);
        l[0] = fminf(l[0], lo.x); l[1] = fminf(l[1], lo.y); l[2] = fminf(l[2], lo.z);
        h[0] = fmaxf(h[0], hi.x); h[1] = fmaxf(h[1], hi.y); h[2] = fmaxf(h[2], hi.z);
    }
    uint32_t q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float u = (pts[3 * (size_t)i + d] - (l[d] - margin)) / (h[d] - l[d] + 2.f * margin);
        q[d] = (uint32_t)fminf(fmaxf(u * 1024.f, 0.f), 1023.f);
    }
    keys[i] = (int32_t)(spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2));
}

struct Best { float d2, d, b1, b2; uint32_t f; };

__device__ __forceinline__ void scan_cluster(const TriRec* __restrict__ tris, uint32_t lo, uint32_t hi, float px, float py, float pz, Best& best) {
    for (uint32_t k = lo; k < hi; ++k) {
        const float4 sp = __ldg(reinterpret_cast<const float4*>(tris + k));         // bounding sphere
        const float dx = px - sp.x, dy = py - sp.y, dz = pz - sp.z;
        const float reach = best.d + sp.w;                                           // |p-c| - r >= best  <=>  |p-c|^2 >= (best+r)^2
        if (dx * dx + dy * dy + dz * dz >= reach * reach) continue;                  // its sphere cannot beat the best
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(tris + k) + 1);
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(tris + k) + 2);
        const float4 q2 = __ldg(reinterpret_cast<const float4*>(tris + k) + 3);
        TriRec t;
        t.ax = q0.x; t.ay = q0.y; t.az = q0.z; t.abx = q1.x; t.aby = q1.y; t.abz = q1.z; t.acx = q2.x; t.acy = q2.y; t.acz = q2.z;
        float b1, b2;
        const float d2 = closest_on_triangle(t, px, py, pz, b1, b2);
        if (d2 < best.d2 || (d2 == best.d2 && k < best.f)) { best.d2 = d2; best.d = sqrtf(d2); best.b1 = b1; best.b2 = b2; best.f = k; }
    }
}

// index of the child box in [lo,hi) nearest to p
__device__ __forceinline__ uint32_t nearest_box(const Box* __restrict__ b, uint32_t lo, uint32_t hi, float px, float py, float pz) {
    uint32_t arg = lo;
    float m = INFINITY;
    for (uint32_t k = lo; k < hi; ++k) {
        const float d = box_dist2(b + k, px, py, pz);
        if (d < m) { m = d; arg = k; }
    }
    return arg;
}

// Exact branch and bound from whatever `best` holds (a real candidate, or just an upper bound with f = none): every box
// that can still hold a closer (or equally close) triangle is visited; `skip_leaf` was scanned by the caller.
__device__ __forceinline__ void search_boxes(const MeshView& m, float px, float py, float pz, uint32_t skip_leaf, Best& best) {
    for (uint32_t s = 0; s < m.n2; ++s) {
        if (box_dist2(m.l2 + s, px, py, pz) > best.d2) continue;
        const uint32_t g_hi = min(m.n1, (s + 1) * kFan);
        for (uint32_t g = s * kFan; g < g_hi; ++g) {
            if (box_dist2(m.l1 + g, px, py, pz) > best.d2) continue;
            const uint32_t c_hi = min(m.n0, (g + 1) * kFan);
            for (uint32_t c = g * kFan; c < c_hi; ++c) {
                if (c == skip_leaf || box_dist2(m.l0 + c, px, py, pz) > best.d2) continue;
                scan_cluster(m.tris, c * kCluster, min(m.n_faces, (c + 1) * kCluster), px, py, pz, best);
            }
        }
    }
}

// Full query: greedy descent to the nearest leaf box for a tight first bound, then the exact search.
// `limit2` < inf: only triangles closer than sqrt(limit2) are of interest (the caller masks everything else out): the bound
// starts there, and a point with nothing inside it comes back with best.f = none after a handful of box tests -- these far
// points are the expensive ones of the unbounded search.  Points inside the limit get the unbounded search's answer (same
// candidate set, same tie rule).
__device__ __forceinline__ void closest_full(const MeshView& m, float px, float py, float pz, Best& best, float limit2 = INFINITY) {
    const uint32_t s0 = nearest_box(m.l2, 0, m.n2, px, py, pz);
    best.d2 = limit2; best.d = limit2 < INFINITY ? sqrtf(limit2) * 1.000001f : INFINITY; best.b1 = 0.f; best.b2 = 0.f; best.f = 0xFFFFFFFFu;
    if (box_dist2(m.l2 + s0, px, py, pz) > limit2) return;                       // nothing of the mesh within the limit
    const uint32_t g0 = nearest_box(m.l1, s0 * kFan, min(m.n1, (s0 + 1) * kFan), px, py, pz);
    const uint32_t c0 = nearest_box(m.l0, g0 * kFan, min(m.n0, (g0 + 1) * kFan), px, py, pz);
    scan_cluster(m.tris, c0 * kCluster, min(m.n_faces, (c0 + 1) * kCluster), px, py, pz, best);
    search_boxes(m, px, py, pz, c0, best);
}

// T_interp = sum_k b_k T[v_k] (utils/ray_utils.py:80), its inverse applied to (p, 1), xyz kept un-normalised (:84); outputs.
__device__ __forceinline__ void warp_outputs(const MeshView& m, const float* __restrict__ T, float threshold, uint32_t i, float px, float py, float pz,
                                             const Best& best, float* __restrict__ can_pts, float* __restrict__ mask, float* __restrict__ closest,
                                             int32_t* __restrict__ face_id, float* __restrict__ dist2_out) {
    const uint32_t best_f = best.f;
    const float bb1 = best.b1, bb2 = best.b2;
    const float bestd2 = best.d2;
    if (best_f == 0xFFFFFFFFu) {         // bounded search, nothing within the limit: masked out, the point itself stands in
        can_pts[3 * (size_t)i] = px; can_pts[3 * (size_t)i + 1] = py; can_pts[3 * (size_t)i + 2] = pz;
        mask[i] = 0.0f;
        if (closest) { closest[3 * (size_t)i] = px; closest[3 * (size_t)i + 1] = py; closest[3 * (size_t)i + 2] = pz; }
        if (face_id) face_id[i] = -1;
        if (dist2_out) dist2_out[i] = bestd2;
        return;
    }
    const TriRec t = m.tris[best_f];
    const float b0 = 1.0f - bb1 - bb2;
    float M[12], c = 0.f;
#pragma unroll
    for (int q = 0; q < 12; ++q) M[q] = 0.f;
    const int32_t vid[3] = {t.i0, t.i1, t.i2};
    const float bw[3] = {b0, bb1, bb2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* Tk = T + 16 * (size_t)vid[k];
#pragma unroll
        for (int q = 0; q < 12; ++q) M[q] = fmaf(bw[k], Tk[q], M[q]);
        c = fmaf(bw[k], Tk[15], c);
    }
    // rows of the 3x4: [m00 m01 m02 tx; m10 m11 m12 ty; m20 m21 m22 tz], last row (0,0,0,c)
    const float m00 = M[0], m01 = M[1], m02 = M[2], tx = M[3];
    const float m10 = M[4], m11 = M[5], m12 = M[6], ty = M[7];
    const float m20 = M[8], m21 = M[9], m22 = M[10], tz = M[11];
    const float c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
    const float det = m00 * c00 + m01 * c01 + m02 * c02;
    const float id = 1.0f / det;
    const float qx = px - tx / c, qy = py - ty / c, qz = pz - tz / c;     // inverse of [[M,t],[0,c]]: M^-1 (p - t/c)
    const float ox = (c00 * qx + (m02 * m21 - m01 * m22) * qy + (m01 * m12 - m02 * m11) * qz) * id;
    const float oy = (c01 * qx + (m00 * m22 - m02 * m20) * qy + (m02 * m10 - m00 * m12) * qz) * id;
    const float oz = (c02 * qx + (m01 * m20 - m00 * m21) * qy + (m00 * m11 - m01 * m10) * qz) * id;
    can_pts[3 * (size_t)i] = ox; can_pts[3 * (size_t)i + 1] = oy; can_pts[3 * (size_t)i + 2] = oz;
    mask[i] = bestd2 < threshold ? 1.0f : 0.0f;
    if (closest) {
        closest[3 * (size_t)i] = t.ax + t.abx * bb1 + t.acx * bb2;
        closest[3 * (size_t)i + 1] = t.ay + t.aby * bb1 + t.acy * bb2;
        closest[3 * (size_t)i + 2] = t.az + t.abz * bb1 + t.acz * bb2;
    }
    if (face_id) face_id[i] = (int32_t)best_f;
    if (dist2_out) dist2_out[i] = bestd2;
}

// pts [n,3] -> can_pts [n,3], mask [n] (dist^2 < threshold), optional closest [n,3], face_id [n], dist2 [n].
// T [n_T,4,4] row-major per-vertex transforms whose last row is (0,0,0,c).
__global__ void __launch_bounds__(256) warp_to_canonical_kernel(const float* __restrict__ pts, const int32_t* __restrict__ order, uint32_t n,
                                                                const MeshView m, const float* __restrict__ T, float threshold,
                                                                float* __restrict__ can_pts, float* __restrict__ mask,
                                                                float* __restrict__ closest, int32_t* __restrict__ face_id,
                                                                float* __restrict__ dist2_out, const float limit2) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t i = order ? (uint32_t)order[slot] : slot;      // spatially sorted queries keep a warp on the same boxes
    const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
    Best best;
    closest_full(m, px, py, pz, best, limit2);
    warp_outputs(m, T, threshold, i, px, py, pz, best, can_pts, mask, closest, face_id, dist2_out);
}

// Samples along a ray are neighbours: a thread walks G consecutive samples of one ray; after the first (full) query each
// next one starts from the triangle-inequality bound dist(p) <= dist(p_prev) + |p - p_prev| and from the leaf that held
// the previous closest triangle, so the exact search only opens the few boxes inside that radius.  Same search, same
// tie rule (smallest triangle index among equals): results are bit-identical to the per-point kernel.
template <int G>
__global__ void __launch_bounds__(256) warp_to_canonical_rays_kernel(const float* __restrict__ pts, uint32_t n_rays, uint32_t n_samples,
                                                                     const MeshView m, const float* __restrict__ T, float threshold,
                                                                     float* __restrict__ can_pts, float* __restrict__ mask,
                                                                     float* __restrict__ closest, int32_t* __restrict__ face_id,
                                                                     float* __restrict__ dist2_out) {
    const uint32_t chunks = (n_samples + G - 1) / G;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rays * chunks) return;
    // chunk-major: the 32 lanes of a warp take the SAME chunk of 32 neighbouring rays (adjacent pixels, same depth range)
    const uint32_t chunk = t / n_rays, ray = t - chunk * n_rays;
    const uint32_t k0 = chunk * G, k1 = min(n_samples, k0 + G);
    Best best;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t i = ray * n_samples + k;
        const float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
        if (k == k0) {
            closest_full(m, px, py, pz, best);
        } else {
            const float dx = px - qx, dy = py - qy, dz = pz - qz;
            const float bound = (best.d + sqrtf(dx * dx + dy * dy + dz * dz)) * 1.000001f + 1e-7f;
            const uint32_t leaf = best.f / kCluster;
            best.d = bound; best.d2 = bound * bound; best.f = 0xFFFFFFFFu; best.b1 = 0.f; best.b2 = 0.f;
            scan_cluster(m.tris, leaf * kCluster, min(m.n_faces, (leaf + 1) * kCluster), px, py, pz, best);
            search_boxes(m, px, py, pz, leaf, best);
        }
        qx = px; qy = py; qz = pz;
        warp_outputs(m, T, threshold, i, px, py, pz, best, can_pts, mask, closest, face_id, dist2_out);
    }
}

// One warp per ray: near = min_v(z0 - dz), far = max_v(z0 + dz) over the vertex spheres the ray pierces
// (utils/ray_utils.py:277-294); rays that miss every sphere fall back to the cube
// (models/instant_nsr.py:147-153 + near_far_from_bound :58-77).
__global__ void __launch_bounds__(256) mesh_near_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t n_rays,
                                                            const float* __restrict__ verts, uint32_t n_verts, float radius, float bound,
                                                            float* __restrict__ near_far) {
    const uint32_t ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= n_rays) return;
    const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
    const float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
    float near = INFINITY, far = -INFINITY;
    const float r2 = radius * radius;
    for (uint32_t v = lane; v < n_verts; v += 32) {
        const float ex = verts[3 * v] - ox, ey = verts[3 * v + 1] - oy, ez = verts[3 * v + 2] - oz;
        const float z0 = ex * dx + ey * dy + ez * dz;
        const float nn = sqrtf(ex * ex + ey * ey + ez * ez);
        const float disc = r2 - (nn * nn - z0 * z0);
        if (disc >= 0.f) {                      // sqrt of a negative -> NaN -> +-inf in the reference
            const float h = sqrtf(disc);
            near = fminf(near, z0 - h);
            far = fmaxf(far, z0 + h);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        near = fminf(near, __shfl_xor_sync(0xffffffffu, near, o));
        far = fmaxf(far, __shfl_xor_sync(0xffffffffu, far, o));
    }
    if (lane == 0) {
        float cn = -INFINITY, cf = INFINITY;
        const float o3[3] = {ox, oy, oz}, d3[3] = {dx, dy, dz};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float den = d3[k] + 1e-15f;
            const float t0 = (-bound - o3[k]) / den, t1 = (bound - o3[k]) / den;
            cn = fmaxf(cn, t0 < t1 ? t0 : t1);
            cf = fminf(cf, t0 > t1 ? t0 : t1);
        }
        cn = fmaxf(cn, 0.05f);
        near_far[2 * ray] = isinf(near) ? cn : near;
        near_far[2 * ray + 1] = isinf(far) ? cf : far;
    }
}

}  // namespace

extern "C" {

uint64_t ac_warp_mesh_bytes(uint32_t n_faces) {
    const MeshView m = mesh_view(nullptr, n_faces);
    return (uint64_t)n_faces * sizeof(TriRec) + (uint64_t)(m.n0 + m.n1 + m.n2) * sizeof(Box);
}

int ac_warp_prepare_mesh(const float* verts, const int32_t* faces, uint32_t face_stride, uint32_t n_faces, void* mesh, void* stream) {
    if (!verts || !faces || !mesh || face_stride < 3) return AC_E_INVALID_ARG;
    if (n_faces == 0) return AC_OK;
    TriRec* tris = reinterpret_cast<TriRec*>(mesh);
    mesh_prepare_kernel<<<(n_faces + 255) / 256, 256, 0, (cudaStream_t)stream>>>(verts, faces, face_stride, n_faces, tris);
    int rc = acb::launched();
    if (rc) return rc;
    const MeshView m = mesh_view(mesh, n_faces);
    mesh_leaf_box_kernel<<<(m.n0 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(tris, n_faces, const_cast<Box*>(m.l0));
    if ((rc = acb::launched())) return rc;
    mesh_inner_box_kernel<<<(m.n1 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m.l0, m.n0, const_cast<Box*>(m.l1));
    if ((rc = acb::launched())) return rc;
    mesh_inner_box_kernel<<<(m.n2 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(m.l1, m.n1, const_cast<Box*>(m.l2));
    return acb::launched();
}

int ac_warp_samples_to_canonical_ordered(const float* pts, const int32_t* order, uint32_t n_pts, const void* mesh, uint32_t n_faces,
                                         const float* T, float threshold, float* can_pts, float* mask, float* closest, int32_t* face_id,
                                         float* dist2, void* stream) {
    if (!pts || !mesh || !T || !can_pts || !mask || n_faces == 0) return AC_E_INVALID_ARG;
    if (n_pts == 0) return AC_OK;
    warp_to_canonical_kernel<<<(n_pts + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pts, order, n_pts, mesh_view(mesh, n_faces), T, threshold,
                                                                                   can_pts, mask, closest, face_id, dist2, INFINITY);
    return acb::launched();
}

int ac_warp_samples_to_canonical_masked(const float* pts, const int32_t* order, uint32_t n_pts, const void* mesh, uint32_t n_faces,
                                        const float* T, float threshold, float* can_pts, float* mask, void* stream) {
    if (!pts || !mesh || !T || !can_pts || !mask || n_faces == 0 || !(threshold > 0.f)) return AC_E_INVALID_ARG;
    if (n_pts == 0) return AC_OK;
    warp_to_canonical_kernel<<<(n_pts + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pts, order, n_pts, mesh_view(mesh, n_faces), T, threshold,
                                                                                   can_pts, mask, nullptr, nullptr, nullptr, threshold);
    return acb::launched();
}

int ac_warp_samples_to_canonical_rays(const float* pts, uint32_t n_rays, uint32_t n_samples, const void* mesh, uint32_t n_faces, const float* T,
                                      float threshold, float* can_pts, float* mask, float* closest, int32_t* face_id, float* dist2, void* stream) {
    if (!pts || !mesh || !T || !can_pts || !mask || n_faces == 0 || n_samples == 0) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    if ((uint64_t)n_rays * n_samples > 0xFFFFFFFFull) return AC_E_INVALID_ARG;
    constexpr int G = 8;
    const uint32_t threads = n_rays * ((n_samples + G - 1) / G);
    warp_to_canonical_rays_kernel<G><<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pts, n_rays, n_samples, mesh_view(mesh, n_faces), T,
                                                                                             threshold, can_pts, mask, closest, face_id, dist2);
    return acb::launched();
}

int ac_warp_samples_to_canonical(const float* pts, uint32_t n_pts, const void* mesh, uint32_t n_faces, const float* T, float threshold,
                                 float* can_pts, float* mask, float* closest, int32_t* face_id, float* dist2, void* stream) {
    return ac_warp_samples_to_canonical_ordered(pts, nullptr, n_pts, mesh, n_faces, T, threshold, can_pts, mask, closest, face_id, dist2, stream);
}

int ac_warp_query_keys(const float* pts, uint32_t n_pts, const void* mesh, uint32_t n_faces, float margin, int32_t* keys, void* stream) {
    if (!pts || !mesh || !keys || n_faces == 0) return AC_E_INVALID_ARG;
    if (n_pts == 0) return AC_OK;
    query_keys_kernel<<<(n_pts + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pts, n_pts, mesh_view(mesh, n_faces), margin, keys);
    return acb::launched();
}

int ac_mesh_guided_near_far(const float* rays_o, const float* rays_d, uint32_t n_rays, const float* verts, uint32_t n_verts,
                            float radius, float bound, float* near_far, void* stream) {
    if (!rays_o || !rays_d || !verts || !near_far) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    mesh_near_far_kernel<<<(n_rays * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays, verts, n_verts, radius, bound,
                                                                                     near_far);
    return acb::launched();
}

}  // extern "C"
