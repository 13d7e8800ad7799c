// frame_ops.cu -- the stages either side of the render core, sm_100a.
//
//   ac_gen_rays            camera -> per-pixel rays on the device (cap2rays / shot_rays, utils/render_utils.py:363-376,
//                          utils/ray_utils.py:25-37; gen_rays_pose, utils/SMPLDataset.py:86-103)
//   ac_select_background   white / black / gaussian grey / blurred chessboard (utils/render_utils.py:953-987)
//   ac_adam_step           torch.optim.Adam (stylize.py:355-363) over ONE flat fp32 buffer, vectorised, with a
//                          read-only fast path for entries that never received a gradient
//   ac_sdf_grid_points     lattice points of extract_fields (models/instant_nsr.py:706-731) for the fused SDF query
//
// All of them are HBM-streaming kernels: 16-byte accesses, grid = a multiple of the SM count, no shared memory.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"

namespace {

struct CamParams {
    double r[9];      // camera-to-world rotation, row-major
    double t[3];      // camera centre
    double fx, fy, cx, cy;
    double x0, xs, y0, ys;   // pixel coordinate of column i = x0 + xs*i (row j likewise)
};

// convention 0 (shot_rays): float64 back-projection of (x, y, 1) to the world, rounded to float32, minus the
// float32 camera centre, normalised in float32 -- the order numpy evaluates it in.
// convention 1 (gen_rays_pose): float32 p = ((x-cx)/fx, -(y-cy)/fy, -1) normalised, rotated by the pose.
__global__ void gen_rays_kernel(const CamParams c, const int W, const int H, const int convention,
                                float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const int n = W * H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int col = i % W, row = i / W;
        float ox, oy, oz, dx, dy, dz;
        if (convention == 0) {
            const double px = (c.x0 + c.xs * col - c.cx) / c.fx, py = (c.y0 + c.ys * row - c.cy) / c.fy;
            const double wx = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, c.r[0]), __dmul_rn(py, c.r[1])), c.r[2]), c.t[0]);
            const double wy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, c.r[3]), __dmul_rn(py, c.r[4])), c.r[5]), c.t[1]);
            const double wz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(px, c.r[6]), __dmul_rn(py, c.r[7])), c.r[8]), c.t[2]);
            ox = (float)c.t[0]; oy = (float)c.t[1]; oz = (float)c.t[2];
            const float vx = __fsub_rn((float)wx, ox), vy = __fsub_rn((float)wy, oy), vz = __fsub_rn((float)wz, oz);
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
            dx = vx / nrm; dy = vy / nrm; dz = vz / nrm;
        } else {
            const float fx = (float)c.fx, fy = (float)c.fy, cx = (float)c.cx, cy = (float)c.cy;
            const float x = (float)(c.x0 + c.xs * col), y = (float)(c.y0 + c.ys * row);
            const float px = __fsub_rn(x, cx) / fx, py = -__fsub_rn(y, cy) / fy, pz = -1.0f;
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), 1.0f));
            const float ux = px / nrm, uy = py / nrm, uz = pz / nrm;
            const float r0 = (float)c.r[0], r1 = (float)c.r[1], r2 = (float)c.r[2], r3 = (float)c.r[3], r4 = (float)c.r[4],
                        r5 = (float)c.r[5], r6 = (float)c.r[6], r7 = (float)c.r[7], r8 = (float)c.r[8];
            dx = __fadd_rn(__fadd_rn(__fmul_rn(ux, r0), __fmul_rn(uy, r1)), __fmul_rn(uz, r2));
            dy = __fadd_rn(__fadd_rn(__fmul_rn(ux, r3), __fmul_rn(uy, r4)), __fmul_rn(uz, r5));
            dz = __fadd_rn(__fadd_rn(__fmul_rn(ux, r6), __fmul_rn(uy, r7)), __fmul_rn(uz, r8));
            ox = (float)c.t[0]; oy = (float)c.t[1]; oz = (float)c.t[2];
        }
        rays_o[3 * i + 0] = ox; rays_o[3 * i + 1] = oy; rays_o[3 * i + 2] = oz;
        rays_d[3 * i + 0] = dx; rays_d[3 * i + 1] = dy; rays_d[3 * i + 2] = dz;
    }
}

// Counter-based generator (two rounds of a 64-bit mix over (seed, index)): independent of launch geometry.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct BlurTaps {
    float kx[5];      // horizontal taps (kernel_size[0] = 5)
    float ky[9];      // vertical taps   (kernel_size[1] = 9)
};

__device__ __forceinline__ int reflect_index(int i, int n) {      // torch 'reflect' padding (no edge repeat)
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// key: 0 white, 1 black, 2 per-ray gaussian grey N(0.5, 0.1) clamped to [0,1], 3 chessboard (0.8 / 0.2 squares of
// side/10 pixels) blurred by a separable 5 x 9 gaussian with reflect padding (torchvision GaussianBlur).
__global__ void background_kernel(const int key, const int n, const int side, const uint64_t seed, const BlurTaps taps,
                                  float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v;
        if (key == 0) v = 1.0f;
        else if (key == 1) v = 0.0f;
        else if (key == 2) {
            const uint64_t h = mix64(mix64(seed) ^ (uint64_t)i);
            const float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);      // (0, 1]
            const float u2 = (float)(uint32_t)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
            const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
            v = fminf(fmaxf(0.5f + 0.1f * g, 0.0f), 1.0f);
        } else {
            const int row = i / side, col = i - row * side;
            const int cell = side / 10 > 0 ? side / 10 : 1;
            float acc = 0.0f;
            for (int a = 0; a < 9; ++a) {
                const int rr = reflect_index(row + a - 4, side);
                float line = 0.0f;
                for (int b = 0; b < 5; ++b) {
                    const int cc = reflect_index(col + b - 2, side);
                    const float board = (((rr / cell) + (cc / cell)) & 1) == 0 ? 0.8f : 0.2f;
                    line = fmaf(taps.kx[b], board, line);
                }
                acc = fmaf(taps.ky[a], line, acc);
            }
            v = acc;
        }
        out[3 * i + 0] = v; out[3 * i + 1] = v; out[3 * i + 2] = v;
    }
}

// torch.optim.Adam, single tensor, no amsgrad, no weight decay, maximize=False (torch/optim/adam.py _single_tensor_adam):
//   m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Entries with g == 0, m == 0 and v == 0 (hash-table slots no ray has ever touched -- most of the fine levels)
// would be rewritten with the same bits: they are read (12 B) and skipped, instead of 16 B read + 12 B written.
struct AdamParams {
    float lr_over_bc1, sqrt_bc2, one_minus_b1, b2, one_minus_b2, eps, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamParams& a) {
    g = g * a.grad_scale;
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), a.one_minus_b1));
    v = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(a.one_minus_b2, g), g));
    const float denom = __fadd_rn(sqrtf(v) / a.sqrt_bc2, a.eps);
    p = __fsub_rn(p, __fmul_rn(a.lr_over_bc1, m / denom));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                                                   float* __restrict__ exp_avg_sq, const uint64_t n, const AdamParams a) {
    const uint64_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 g = g4[i];
        float4 m = m4[i], v = v4[i];
        const bool idle = g.x == 0.f && g.y == 0.f && g.z == 0.f && g.w == 0.f && m.x == 0.f && m.y == 0.f && m.z == 0.f && m.w == 0.f &&
                          v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f;
        if (idle) continue;
        float4 p = p4[i];
        adam_one(p.x, g.x, m.x, v.x, a); adam_one(p.y, g.y, m.y, v.y, a);
        adam_one(p.z, g.z, m.z, v.z, a); adam_one(p.w, g.w, m.w, v.w, a);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {             // tail (n not a multiple of 4)
        const uint64_t i = n4 * 4 + threadIdx.x;
        float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
        adam_one(p, grad[i], m, v, a);
        param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
    }
}

// extract_fields (models/instant_nsr.py:706-731): lattice of `res`^3 points, X = linspace(lo, hi, res) per axis,
// point index (i*res + j)*res + k <-> (X[i], Y[j], Z[k]) -- the layout of the reference's `u[xi, yi, zi]`.
// Block [i0, i0+ni) of the slowest axis, so the volume can be produced slab by slab.
__global__ void grid_points_kernel(const float lo0, const float lo1, const float lo2, const float hi0, const float hi1, const float hi2,
                                   const int res, const int i0, const int ni, float* __restrict__ pts) {
    const size_t n = (size_t)ni * res * res;
    const float step0 = (hi0 - lo0) / (float)(res - 1), step1 = (hi1 - lo1) / (float)(res - 1), step2 = (hi2 - lo2) / (float)(res - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % res), j = (int)((t / res) % res), i = i0 + (int)(t / ((size_t)res * res));
        // torch.linspace: lower half start + step*i, upper half end - step*(n-1-i)
        const float x = i < res / 2 ? fmaf(step0, (float)i, lo0) : fmaf(-step0, (float)(res - 1 - i), hi0);
        const float y = j < res / 2 ? fmaf(step1, (float)j, lo1) : fmaf(-step1, (float)(res - 1 - j), hi1);
        const float z = k < res / 2 ? fmaf(step2, (float)k, lo2) : fmaf(-step2, (float)(res - 1 - k), hi2);
        pts[3 * t + 0] = x; pts[3 * t + 1] = y; pts[3 * t + 2] = z;
    }
}

// Iso-surface of a lattice volume by marching tetrahedra (Kuhn split of every cell into the 6 tetrahedra around the
// 0-7 diagonal: face diagonals agree between neighbouring cells, so the surface is watertight).  One thread per cell;
// triangles are appended through one atomic counter as 3 x (position, lattice-edge key); the host welds vertices by key.
// Stands in for `mcubes.marching_cubes(-sdf, 0)` (models/instant_nsr.py:745-752; PyMCubes is not part of the image).
// Triangles are oriented so that their normal points towards increasing field value (out of the surface for an SDF).
struct IsoParams {
    float lo[3], hi[3];
    int res;
    float threshold;
    unsigned long long capacity;      // triangles that fit
};

__global__ void __launch_bounds__(256) iso_tets_kernel(const float* __restrict__ vol, const IsoParams q, float* __restrict__ tri_pos,
                                                       long long* __restrict__ tri_key, unsigned long long* __restrict__ counter) {
    const int R = q.res, C = R - 1;
    const size_t cells = (size_t)C * C * C;
    const float sx = (q.hi[0] - q.lo[0]) / (float)(R - 1), sy = (q.hi[1] - q.lo[1]) / (float)(R - 1), sz = (q.hi[2] - q.lo[2]) / (float)(R - 1);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < cells; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % C), j = (int)((t / C) % C), i = (int)(t / ((size_t)C * C));
        float v[8];
        long long id[8];
        bool any_in = false, any_out = false;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ci = i + (c & 1), cj = j + ((c >> 1) & 1), ck = k + ((c >> 2) & 1);
            id[c] = ((long long)ci * R + cj) * R + ck;
            v[c] = vol[id[c]];
            any_in |= v[c] < q.threshold;
            any_out |= !(v[c] < q.threshold);
        }
        if (!(any_in && any_out)) continue;
        const int tets[6][4] = {{0, 1, 3, 7}, {0, 3, 2, 7}, {0, 2, 6, 7}, {0, 6, 4, 7}, {0, 4, 5, 7}, {0, 5, 1, 7}};
        for (int tt = 0; tt < 6; ++tt) {
            int in[4], out[4], n_in = 0, n_out = 0;
            for (int a = 0; a < 4; ++a) {
                const int c = tets[tt][a];
                if (v[c] < q.threshold) in[n_in++] = c; else out[n_out++] = c;
            }
            if (n_in == 0 || n_in == 4) continue;
            int ea[4], eb[4], n_e;                       // crossing edges: (inside corner, outside corner)
            if (n_in == 1) { n_e = 3; for (int a = 0; a < 3; ++a) { ea[a] = in[0]; eb[a] = out[a]; } }
            else if (n_in == 3) { n_e = 3; for (int a = 0; a < 3; ++a) { ea[a] = in[a]; eb[a] = out[0]; } }
            else { n_e = 4; ea[0] = in[0]; eb[0] = out[0]; ea[1] = in[0]; eb[1] = out[1]; ea[2] = in[1]; eb[2] = out[1]; ea[3] = in[1]; eb[3] = out[0]; }
            float P[4][3];
            long long key[4];
            float cin[3] = {0.f, 0.f, 0.f}, cout[3] = {0.f, 0.f, 0.f};
            for (int a = 0; a < n_e; ++a) {
                const int p = ea[a], r = eb[a];
                // interpolate from the lower lattice id so both cells sharing the edge produce identical bits
                const int lo_c = id[p] < id[r] ? p : r, hi_c = id[p] < id[r] ? r : p;
                const float w = (q.threshold - v[lo_c]) / (v[hi_c] - v[lo_c]);
                const float ax = (float)(i + (lo_c & 1)), ay = (float)(j + ((lo_c >> 1) & 1)), az = (float)(k + ((lo_c >> 2) & 1));
                const float bx = (float)(i + (hi_c & 1)), by = (float)(j + ((hi_c >> 1) & 1)), bz = (float)(k + ((hi_c >> 2) & 1));
                P[a][0] = q.lo[0] + sx * (ax + w * (bx - ax));
                P[a][1] = q.lo[1] + sy * (ay + w * (by - ay));
                P[a][2] = q.lo[2] + sz * (az + w * (bz - az));
                key[a] = (id[lo_c] << 32) | (id[hi_c] & 0xFFFFFFFFll);
            }
            for (int a = 0; a < n_in; ++a) { cin[0] += (float)(in[a] & 1) / n_in; cin[1] += (float)((in[a] >> 1) & 1) / n_in; cin[2] += (float)((in[a] >> 2) & 1) / n_in; }
            for (int a = 0; a < n_out; ++a) { cout[0] += (float)(out[a] & 1) / n_out; cout[1] += (float)((out[a] >> 1) & 1) / n_out; cout[2] += (float)((out[a] >> 2) & 1) / n_out; }
            const float gx = (cout[0] - cin[0]) * sx, gy = (cout[1] - cin[1]) * sy, gz = (cout[2] - cin[2]) * sz;     // inside -> outside
            const int n_tri = n_e - 2;
            for (int f = 0; f < n_tri; ++f) {
                int a = 0, b = f + 1, c = f + 2;
                const float ux = P[b][0] - P[a][0], uy = P[b][1] - P[a][1], uz = P[b][2] - P[a][2];
                const float wx = P[c][0] - P[a][0], wy = P[c][1] - P[a][1], wz = P[c][2] - P[a][2];
                const float nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
                if (nx * gx + ny * gy + nz * gz < 0.f) { const int s = b; b = c; c = s; }
                const unsigned long long slot = atomicAdd(counter, 1ull);
                if (slot >= q.capacity) continue;                       // counted, not stored: the caller re-runs with room
                const int order[3] = {a, b, c};
                for (int e = 0; e < 3; ++e) {
                    tri_pos[(slot * 3 + e) * 3 + 0] = P[order[e]][0];
                    tri_pos[(slot * 3 + e) * 3 + 1] = P[order[e]][1];
                    tri_pos[(slot * 3 + e) * 3 + 2] = P[order[e]][2];
                    tri_key[slot * 3 + e] = key[order[e]];
                }
            }
        }
    }
}

inline int grid_for(uint64_t work, int block, int per_sm) {
    const uint64_t want = (work + block - 1) / block;
    const uint64_t cap = (uint64_t)acb::sm_count() * per_sm;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

extern "C" {

int ac_gen_rays(const double* c2w, double fx, double fy, double cx, double cy, uint32_t W, uint32_t H, double x0, double x_step,
                double y0, double y_step, int convention, float* rays_o, float* rays_d, void* stream) {
    if (!c2w || !rays_o || !rays_d || W == 0 || H == 0 || (convention != 0 && convention != 1)) return AC_E_INVALID_ARG;
    if ((uint64_t)W * H > 0x7FFFFFFFull) return AC_E_INVALID_ARG;
    CamParams c;
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) c.r[3 * r + k] = c2w[4 * r + k];
        c.t[r] = c2w[4 * r + 3];
    }
    c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy; c.x0 = x0; c.xs = x_step; c.y0 = y0; c.ys = y_step;
    gen_rays_kernel<<<grid_for((uint64_t)W * H, 256, 8), 256, 0, (cudaStream_t)stream>>>(c, (int)W, (int)H, convention, rays_o, rays_d);
    return acb::launched();
}

int ac_select_background(int key, uint32_t n_rays, uint64_t seed, float sigma, float* out, void* stream) {
    if (!out || n_rays == 0 || n_rays > 0x7FFFFFFFu) return AC_E_INVALID_ARG;
    key = ((key % 4) + 4) % 4;
    BlurTaps taps;
    int side = 0;
    if (key == 3) {
        side = (int)floor(sqrt((double)n_rays));
        if ((uint64_t)side * side != n_rays || !(sigma > 0.0f)) return AC_E_INVALID_ARG;     // "assume sqrt is integer" (:973)
        // torchvision _get_gaussian_kernel1d: x = linspace(-(k-1)/2, (k-1)/2, k); pdf = exp(-0.5 (x/sigma)^2); pdf / sum
        float sx = 0.f, sy = 0.f;
        for (int b = 0; b < 5; ++b) { const float x = (float)(b - 2) / sigma; taps.kx[b] = expf(-0.5f * x * x); sx += taps.kx[b]; }
        for (int a = 0; a < 9; ++a) { const float x = (float)(a - 4) / sigma; taps.ky[a] = expf(-0.5f * x * x); sy += taps.ky[a]; }
        for (int b = 0; b < 5; ++b) taps.kx[b] /= sx;
        for (int a = 0; a < 9; ++a) taps.ky[a] /= sy;
    } else {
        for (int b = 0; b < 5; ++b) taps.kx[b] = 0.f;
        for (int a = 0; a < 9; ++a) taps.ky[a] = 0.f;
    }
    background_kernel<<<grid_for(n_rays, 256, 8), 256, 0, (cudaStream_t)stream>>>(key, (int)n_rays, side, seed, taps, out);
    return acb::launched();
}

int ac_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, uint64_t n, float lr, float beta1,
                 float beta2, float eps, uint32_t step, float grad_scale, void* stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || step == 0) return AC_E_INVALID_ARG;
    if (n == 0) return AC_OK;
    if ((((uintptr_t)param) | ((uintptr_t)grad) | ((uintptr_t)exp_avg) | ((uintptr_t)exp_avg_sq)) & 15u) return AC_E_INVALID_ARG;
    AdamParams a;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    a.lr_over_bc1 = (float)((double)lr / bc1);
    a.sqrt_bc2 = (float)sqrt(bc2);
    a.one_minus_b1 = 1.0f - beta1; a.b2 = beta2; a.one_minus_b2 = 1.0f - beta2; a.eps = eps; a.grad_scale = grad_scale;
    adam_kernel<<<grid_for(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, a);
    return acb::launched();
}

int ac_sdf_grid_points(const float* bound_min, const float* bound_max, uint32_t resolution, uint32_t i0, uint32_t ni, float* pts,
                       void* stream) {
    if (!bound_min || !bound_max || !pts || resolution < 2 || ni == 0 || i0 + ni > resolution) return AC_E_INVALID_ARG;
    grid_points_kernel<<<grid_for((uint64_t)ni * resolution * resolution, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        bound_min[0], bound_min[1], bound_min[2], bound_max[0], bound_max[1], bound_max[2], (int)resolution, (int)i0, (int)ni, pts);
    return acb::launched();
}

int ac_iso_surface(const float* volume, const float* bound_min, const float* bound_max, uint32_t resolution, float threshold,
                   float* tri_pos, int64_t* tri_key, uint64_t capacity, uint64_t* counter, void* stream) {
    if (!volume || !bound_min || !bound_max || !counter || resolution < 2 || resolution > 1625) return AC_E_INVALID_ARG;   // ids fit 32 bits
    if (capacity > 0 && (!tri_pos || !tri_key)) return AC_E_INVALID_ARG;
    IsoParams q;
    for (int a = 0; a < 3; ++a) { q.lo[a] = bound_min[a]; q.hi[a] = bound_max[a]; }
    q.res = (int)resolution; q.threshold = threshold; q.capacity = capacity;
    const uint64_t cells = (uint64_t)(resolution - 1) * (resolution - 1) * (resolution - 1);
    iso_tets_kernel<<<grid_for(cells, 256, 8), 256, 0, (cudaStream_t)stream>>>(volume, q, tri_pos, reinterpret_cast<long long*>(tri_key),
                                                                                 reinterpret_cast<unsigned long long*>(counter));
    return acb::launched();
}

}  // extern "C"
