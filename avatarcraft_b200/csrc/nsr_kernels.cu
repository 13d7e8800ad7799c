// nsr_kernels.cu -- fused Instant-NSR render core for sm_100a and its C ABI.
//
// Execution model: ONE WARP OWNS ONE RAY from the box intersection to the composited pixel.
// All per-ray state (sorted depths, their SDF, CDF scratch) lives in a 2 KB shared-memory
// slice private to the warp, so the kernel needs a single __syncthreads (after the MLP
// weights and level table are staged) and otherwise only __syncwarp / shuffles.  The 128
// section samples of a ray are visited 32 at a time in depth order, which lets the NeuS
// transmittance be carried across iterations with a 5-step warp-shuffle product scan.
// Compulsory HBM traffic is 24 B in + 32 B out per ray; the hash table (49 MB) is read
// through L2/L1 with 8 B gathers.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"

using namespace acb;

namespace {

constexpr int kWarps = 8;                 // rays per CTA
constexpr int kMaxT = 128;

struct RenderParams {
    const float2* table;
    const int32_t* offsets;
    const float* blob;
    const float* variance;
    float S;
    uint32_t H;
    ac_nsr_render_args a;
    float* eik_partial;   // [n_rays][2]
};

__device__ __forceinline__ void stage_model(const float* __restrict__ blob, const int32_t* __restrict__ offsets,
                                            float S, uint32_t H, float* sw, LevelMeta* lv) {
    const float4* src = reinterpret_cast<const float4*>(blob);
    float4* dst = reinterpret_cast<float4*>(sw);
    for (int i = threadIdx.x; i < BLOB_FLOATS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(offsets, threadIdx.x, S, H, 3);
    __syncthreads();
}

constexpr size_t kStageBytes = BLOB_FLOATS * sizeof(float) + kLevels * sizeof(LevelMeta);

// SDF network weights (W0 | b0 | W1T | b1 = the first OFF_C0 floats of the blob) in the constant bank for the
// point-query kernels of the training path: with the hidden-unit loop fully unrolled every weight is an
// immediate constant-bank operand of its FFMA -- no shared-memory loads (they were ~3000 LDS per point in the
// backward kernel).  Refreshed by a 13.6 KB D2D copy on the launch stream.
__constant__ float c_sdf[OFF_C0];

__device__ __forceinline__ void stage_levels(const int32_t* __restrict__ offsets, float S, uint32_t H, LevelMeta* lv) {
    if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(offsets, threadIdx.x, S, H, 3);
    __syncthreads();
}
int refresh_c_sdf(const float* blob, cudaStream_t st) {
    return cudaMemcpyToSymbolAsync(c_sdf, blob, OFF_C0 * sizeof(float), 0, cudaMemcpyDeviceToDevice, st) == cudaSuccess ? AC_OK : acb::cuda_fail();
}

__global__ void __launch_bounds__(kWarps * 32) nsr_render_kernel(const RenderParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sw = reinterpret_cast<float*>(smem_raw);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(sw + BLOB_FLOATS);
    float* rows = reinterpret_cast<float*>(lv + kLevels);
    stage_model(p.blob, p.offsets, p.S, p.H, sw, lv);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ray = blockIdx.x * kWarps + warp;
    if (ray >= p.a.n_rays) return;
    float* zs = rows + warp * 4 * kMaxT;
    float* sdfs = zs + kMaxT;
    float* ta = sdfs + kMaxT;
    float* tb = ta + kMaxT;

    const float bound = p.a.bound;
    const float2* __restrict__ table = p.table;
    Ray r;
    r.ox = p.a.rays_o[3 * ray + 0]; r.oy = p.a.rays_o[3 * ray + 1]; r.oz = p.a.rays_o[3 * ray + 2];
    r.dx = p.a.rays_d[3 * ray + 0]; r.dy = p.a.rays_d[3 * ray + 1]; r.dz = p.a.rays_d[3 * ray + 2];
    float near, far;
    ray_box(r, bound, near, far);
    const int N0 = (int)p.a.num_steps;
    const float span = far - near;
    const float sample_dist = span / (float)N0;               // keeps the COARSE count (:160)

    // ---- coarse samples (:155-174) and their SDF (:178) ----
    int T = N0;
    for (int k = lane; k < N0; k += 32) {
        float z = near + span * linspace01(k, N0);
        if (p.a.jitter) z = z + (p.a.jitter[(size_t)ray * N0 + k] - 0.5f) * sample_dist;
        zs[k] = z;
        if (p.a.upsample_steps > 0) {
            float x, y, zz;
            ray_point(r, z, x, y, zz);
            float o[1];
            sdf_point<false>(table, lv, sw, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
                             clampf(zz, -bound, bound), o);
            sdfs[k] = o[0];
        }
    }
    __syncwarp();

    // ---- importance rounds (:182-184) ----
    const int rounds = (int)p.a.upsample_steps / 16;
    for (int i = 0; i < rounds; ++i) {
        float z_new; int below, above;
        importance_round(r, zs, sdfs, ta, tb, T, (float)(64 << i), lane, z_new, below, above);
        float s_new = 0.0f;
        const bool last = (i + 1 == rounds);
        if (!last && lane < 16) {
            float x, y, zz;
            ray_point(r, z_new, x, y, zz);
            float o[1];
            sdf_point<false>(table, lv, sw, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
                             clampf(zz, -bound, bound), o);
            s_new = o[0];
        }
        int pos_old[4], pos_new;
        merge_positions(zs, T, z_new, lane, pos_old, pos_new);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = lane + 32 * q;
            if (k < T) { ta[pos_old[q]] = zs[k]; tb[pos_old[q]] = sdfs[k]; }
        }
        if (lane < 16) { ta[pos_new] = z_new; tb[pos_new] = s_new; }
        __syncwarp();
        float* t0 = zs; zs = ta; ta = t0;
        float* t1 = sdfs; sdfs = tb; tb = t1;
        T += 16;
    }

    // ---- render core (:186-299): 32 section samples per iteration, in depth order ----
    const float inv_s = clampf(expf(p.variance[0] * 10.0f), 1e-6f, 1e6f);
    const float eps = 0.005f * (1.0f - p.a.normal_epsilon_ratio);
    const float car = p.a.cos_anneal_ratio;
    float carry = 1.0f;                    // transmittance entering this block of 32 samples
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_nx = 0.f, acc_ny = 0.f, acc_nz = 0.f;
    float acc_w = 0.f, acc_d = 0.f, eik_num = 0.f, eik_den = 0.f;
    for (int k0 = 0; k0 < T; k0 += 32) {
        const int k = k0 + lane;
        const bool live = k < T;
        float alpha = 0.f, col[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f}, zk = 0.f;
        if (live) {
            zk = zs[k];
            const float delta = k < T - 1 ? zs[k + 1] - zk : sample_dist;
            const float zmid = k < T - 1 ? zk + 0.5f * delta : zk;
            float px, py, pz;
            ray_point(r, zmid, px, py, pz);
            px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
            float o16[16];
            sdf_point<true>(table, lv, sw, bound, px, py, pz, o16);
            // central differences (:687-704); every shifted point is re-clamped
            float g[3];
#pragma unroll 1
            for (int ax = 0; ax < 3; ++ax) {
                float f2[2];
#pragma unroll 1
                for (int sg = 0; sg < 2; ++sg) {
                    const float e = sg == 0 ? eps : -eps;
                    const float qx = ax == 0 ? clampf(px + e, -bound, bound) : px;
                    const float qy = ax == 1 ? clampf(py + e, -bound, bound) : py;
                    const float qz = ax == 2 ? clampf(pz + e, -bound, bound) : pz;
                    float o[1];
                    sdf_point<false>(table, lv, sw, bound, qx, qy, qz, o);
                    f2[sg] = o[0];
                }
                g[ax] = 0.5f * (f2[0] - f2[1]) / eps;
            }
            const float gn = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            const float inv = 1e-5f + gn;
            nrm[0] = g[0] / inv; nrm[1] = g[1] / inv; nrm[2] = g[2] / inv;
            float cin[kColInPad];
            cin[0] = px; cin[1] = py; cin[2] = pz; cin[3] = nrm[0]; cin[4] = nrm[1]; cin[5] = nrm[2];
#pragma unroll
            for (int q = 0; q < 15; ++q) cin[6 + q] = o16[1 + q];
            cin[21] = cin[22] = cin[23] = 0.f;
            color_mlp(sw, cin, col);
            // NeuS alpha (:221-243)
            const float cosv = r.dx * nrm[0] + r.dy * nrm[1] + r.dz * nrm[2];
            const float it = -(softplus100(-cosv * 0.5f + 0.5f) * (1.0f - car) + softplus100(-cosv) * car);
            const float hs = it * delta * 0.5f;
            const float c0 = sigmoidf((o16[0] - hs) * inv_s), c1 = sigmoidf((o16[0] + hs) * inv_s);
            alpha = clampf((c0 - c1 + 1e-5f) / (c0 + 1e-5f), 0.0f, 1.0f);
            if (p.a.alpha_mask) alpha = alpha * p.a.alpha_mask[(size_t)ray * T + k];
            // eikonal terms (:265-272)
            const float pn = sqrtf(px * px + py * py + pz * pz);
            if (pn < 1.2f) { eik_num += (gn - 1.0f) * (gn - 1.0f); eik_den += 1.0f; }
        }
        float blk;
        const float tr = warp_excl_prod(live ? (1.0f - alpha + 1e-7f) : 1.0f, lane, blk) * carry;
        carry *= blk;
        const float w = alpha * tr;
        if (live) {
            acc_r += w * col[0]; acc_g += w * col[1]; acc_b += w * col[2];
            acc_nx += w * nrm[0]; acc_ny += w * nrm[1]; acc_nz += w * nrm[2];
            acc_w += w;
            acc_d += w * clampf((zk - near) / span, 0.0f, 1.0f);
            const size_t s = (size_t)ray * T + k;
            if (p.a.weights) p.a.weights[s] = w;
            if (p.a.pts_alpha) p.a.pts_alpha[s] = alpha;
            if (p.a.z_vals) p.a.z_vals[s] = zk;
            if (p.a.pts_color) { p.a.pts_color[3 * s] = col[0]; p.a.pts_color[3 * s + 1] = col[1]; p.a.pts_color[3 * s + 2] = col[2]; }
        }
    }
    acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
    acc_nx = warp_sum(acc_nx); acc_ny = warp_sum(acc_ny); acc_nz = warp_sum(acc_nz);
    acc_w = warp_sum(acc_w); acc_d = warp_sum(acc_d);
    eik_num = warp_sum(eik_num); eik_den = warp_sum(eik_den);
    if (lane == 0) {
        float bg[3] = {1.f, 1.f, 1.f};
        if (p.a.bg_color) { bg[0] = p.a.bg_color[3 * ray]; bg[1] = p.a.bg_color[3 * ray + 1]; bg[2] = p.a.bg_color[3 * ray + 2]; }
        const float rest = 1.0f - acc_w;
        p.a.rgb[3 * ray + 0] = acc_r + rest * bg[0];
        p.a.rgb[3 * ray + 1] = acc_g + rest * bg[1];
        p.a.rgb[3 * ray + 2] = acc_b + rest * bg[2];
        p.a.depth[ray] = acc_d;
        p.a.weight_sum[ray] = acc_w;
        p.a.normal[3 * ray + 0] = acc_nx; p.a.normal[3 * ray + 1] = acc_ny; p.a.normal[3 * ray + 2] = acc_nz;
        p.eik_partial[2 * ray + 0] = eik_num;
        p.eik_partial[2 * ray + 1] = eik_den;
    }
}

// Deterministic reduction of the per-ray eikonal partials (:270-272), one block per segment.
__global__ void __launch_bounds__(1024) eikonal_reduce_kernel(const float* __restrict__ partial, uint32_t n, uint32_t seg, float* out) {
    __shared__ double s_num[32], s_den[32];
    double num = 0.0, den = 0.0;
    const uint32_t lo = blockIdx.x * seg, hi = min(n, lo + seg);
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) { num += partial[2 * i]; den += partial[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
    if ((threadIdx.x & 31) == 0) { s_num[threadIdx.x >> 5] = num; s_den[threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x < 32) {
        num = s_num[threadIdx.x]; den = s_den[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
        if (threadIdx.x == 0) out[blockIdx.x] = (float)num / ((float)den + 1e-5f);
    }
}

// ---- flat-point queries (NeRFNetwork.forward_sdf / forward_color / gradient) ----
__global__ void __launch_bounds__(256) forward_sdf_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                          float S, uint32_t H, const float* __restrict__ x, float* __restrict__ out,
                                                          uint32_t B, float bound) {
    __shared__ LevelMeta lv[kLevels];
    stage_levels(offsets, S, H, lv);
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float in[kSdfInPad], o16[16];
    encode_point(table, lv, bound, x[3 * (size_t)b], x[3 * (size_t)b + 1], x[3 * (size_t)b + 2], in);
#pragma unroll
    for (int o = 0; o < 16; ++o) o16[o] = c_sdf[OFF_B1 + o];
#pragma unroll
    for (int j = 0; j < kHidden; ++j) {
        float a = c_sdf[OFF_B0 + j];
#pragma unroll
        for (int k = 0; k < 35; ++k) a = fmaf(c_sdf[OFF_W0 + j * kSdfInPad + k], in[k], a);
        const float h = softplus100(a);
#pragma unroll
        for (int o = 0; o < 16; ++o) o16[o] = fmaf(c_sdf[OFF_W1T + j * 16 + o], h, o16[o]);
    }
    float4* dst = reinterpret_cast<float4*>(out + 16 * (size_t)b);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
}

__global__ void __launch_bounds__(256) fd_gradient_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                          const float* __restrict__ blob, float S, uint32_t H,
                                                          const float* __restrict__ x, float* __restrict__ grad, uint32_t B,
                                                          float bound, float eps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sw = reinterpret_cast<float*>(smem_raw);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(sw + BLOB_FLOATS);
    stage_model(blob, offsets, S, H, sw, lv);
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float px = x[3 * b], py = x[3 * b + 1], pz = x[3 * b + 2];
#pragma unroll 1
        for (int ax = 0; ax < 3; ++ax) {
            float f2[2];
#pragma unroll 1
            for (int sg = 0; sg < 2; ++sg) {
                const float e = sg == 0 ? eps : -eps;
                float o[1];
                sdf_point<false>(table, lv, sw, bound, clampf(ax == 0 ? px + e : px, -bound, bound),
                                 clampf(ax == 1 ? py + e : py, -bound, bound), clampf(ax == 2 ? pz + e : pz, -bound, bound), o);
                f2[sg] = o[0];
            }
            grad[3 * (size_t)b + ax] = 0.5f * (f2[0] - f2[1]) / eps;
        }
    }
}

__global__ void __launch_bounds__(256) forward_color_kernel(const float* __restrict__ blob, const float* __restrict__ x,
                                                            const float* __restrict__ n, const float* __restrict__ feat,
                                                            float* __restrict__ rgb, uint32_t B, const float* __restrict__ bias0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sw = reinterpret_cast<float*>(smem_raw);
    const float4* src = reinterpret_cast<const float4*>(blob);
    for (int i = threadIdx.x; i < BLOB_FLOATS / 4; i += blockDim.x) reinterpret_cast<float4*>(sw)[i] = __ldg(src + i);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;   // one point per thread: a grid-stride loop here
    if (b >= B) return;                                         // lets nvcc hoist the weight rows into registers
    float cin[kColInPad], c[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { cin[q] = x[3 * (size_t)b + q]; cin[3 + q] = n[3 * (size_t)b + q]; cin[21 + q] = 0.f; }
#pragma unroll
    for (int q = 0; q < 15; ++q) cin[6 + q] = feat[15 * (size_t)b + q];
    color_mlp(sw, cin, c, bias0 ? bias0 + 64 * (size_t)b : nullptr);
    rgb[3 * (size_t)b] = c[0]; rgb[3 * (size_t)b + 1] = c[1]; rgb[3 * (size_t)b + 2] = c[2];
}

// out [n,64] = sh [n,16] w_sh[64,16]^T: the ray direction's contribution to colour layer 0 (use_viewdirs).
__global__ void __launch_bounds__(256) viewdir_bias_kernel(const float* __restrict__ sh, const float* __restrict__ w_sh, uint32_t n,
                                                           float* __restrict__ out) {
    __shared__ float w[64 * 16];
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) w[i] = w_sh[i];
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * 64) return;
    const size_t ray = t >> 6;
    const int j = (int)(t & 63);
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) a = fmaf(w[j * 16 + k], sh[ray * 16 + k], a);
    out[t] = a;
}


// ---- training: fused backward of forward_sdf for a flat list of points -----------------------
// Recomputes hash features and the hidden layer, back-propagates grad_out [B,16] to
//   * the hash table: w * d(feature) scattered into the 8 corners of 16 levels with fp32
//     red.global.add.v2 (same arithmetic as kernel_grid_backward, hashencoder.cu:223-308),
//   * per-point layer deltas that the host turns into weight gradients with plain GEMMs:
//       delta_a [64,B] = dL/d(pre-activation), hidden [64,B] = softplus output, feats [36,B] = the layer input (xyz | features | 1)
//     (fp32: bf16 operands were measured -- 1.8 ms faster per 4096-ray patch, but 2 % error on the bias gradients, whose terms cancel)
//     (unit-major: consecutive points are consecutive addresses, so every store of a warp is one 128 B line; the
//     point-major layout cost 32 sectors per store instruction and most of this kernel's time).
// softplus'(a) = sigmoid(100 a) (1 above torch's threshold 100a > 20).
__global__ void __launch_bounds__(256) sdf_backward_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                           const float* __restrict__ blob, float S, uint32_t H,
                                                           const float* __restrict__ x, const float* __restrict__ gout,
                                                           uint32_t B, float bound, float* __restrict__ grad_table,
                                                           float* __restrict__ delta_a, float* __restrict__ hidden,
                                                           float* __restrict__ feats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sw = reinterpret_cast<float*>(smem_raw);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(sw + BLOB_FLOATS);
    float* w0f = reinterpret_cast<float*>(lv + kLevels);          // [64][32] feature columns of W0, 16 B aligned rows
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) w0f[i] = __ldg(blob + OFF_W0 + (i >> 5) * kSdfInPad + 3 + (i & 31));
    stage_model(blob, offsets, S, H, sw, lv);
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {     // persistent: weights staged once
        const float px = x[3 * (size_t)b], py = x[3 * (size_t)b + 1], pz = x[3 * (size_t)b + 2];
        float in[kSdfInPad];
        encode_point(table, lv, bound, px, py, pz, in);
        float g[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(gout + 16 * (size_t)b + 4 * q);
            g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
        }
        float din[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) din[k] = 0.f;
#pragma unroll 1
        for (int j = 0; j < kHidden; ++j) {
            const float4* __restrict__ wr = reinterpret_cast<const float4*>(sw + OFF_W0 + j * kSdfInPad);
            float a = sw[OFF_B0 + j];
#pragma unroll
            for (int q = 0; q < kSdfInPad / 4; ++q) {
                const float4 w4 = wr[q];
                a = fmaf(w4.x, in[4 * q + 0], a); a = fmaf(w4.y, in[4 * q + 1], a);
                a = fmaf(w4.z, in[4 * q + 2], a); a = fmaf(w4.w, in[4 * q + 3], a);
            }
            const float h = softplus100(a);
            const float4* __restrict__ w1 = reinterpret_cast<const float4*>(sw + OFF_W1T + j * 16);
            float dh = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = w1[q];
                dh = fmaf(w4.x, g[4 * q], dh); dh = fmaf(w4.y, g[4 * q + 1], dh);
                dh = fmaf(w4.z, g[4 * q + 2], dh); dh = fmaf(w4.w, g[4 * q + 3], dh);
            }
            const float da = a * 100.0f > 20.0f ? dh : dh * sigmoidf(a * 100.0f);
            delta_a[(size_t)j * B + b] = da;            // unit-major [64][B]: a warp's 32 points write one 128 B line
            hidden[(size_t)j * B + b] = h;
            const float4* __restrict__ wf = reinterpret_cast<const float4*>(w0f + j * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w4 = wf[q];
                din[4 * q + 0] = fmaf(w4.x, da, din[4 * q + 0]); din[4 * q + 1] = fmaf(w4.y, da, din[4 * q + 1]);
                din[4 * q + 2] = fmaf(w4.z, da, din[4 * q + 2]); din[4 * q + 3] = fmaf(w4.w, da, din[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 35; ++q) feats[(size_t)q * B + b] = in[q];         // input-major [36][B] = (x, y, z, 32 features, 1), coalesced
        feats[(size_t)35 * B + b] = 1.0f;                                      // ones row: the bias gradient falls out of the same GEMM
        // scatter into the table
        const float two_b = 2.0f * bound;
        const float u = (px + bound) / two_b, v = (py + bound) / two_b, w = (pz + bound) / two_b;
        if ((u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f)) continue;
#pragma unroll
        for (int l = 0; l < kLevels; ++l) {
            const LevelMeta m = lv[l];
            float fx = fmaf(u, m.scale, 0.5f), fy = fmaf(v, m.scale, 0.5f), fz = fmaf(w, m.scale, 0.5f);
            const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
            const uint32_t ix = (uint32_t)flx, iy = (uint32_t)fly, iz = (uint32_t)flz;
            fx -= flx; fy -= fly; fz -= flz;
            float2* __restrict__ dst = reinterpret_cast<float2*>(grad_table) + m.offset;
            const float gx = din[2 * l], gy = din[2 * l + 1];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t cx = ix + (k & 1), cy = iy + ((k >> 1) & 1), cz = iz + ((k >> 2) & 1);
                uint32_t slot;
                if (m.hashed == 0u) slot = cx + cy * m.res1 + cz * m.res1 * m.res1;
                else slot = wrap_slot(cx ^ (cy * 2654435761u) ^ (cz * 805459861u), m);
                const float wgt = (((k & 1) ? fx : 1.0f - fx) * ((k & 2) ? fy : 1.0f - fy)) * ((k & 4) ? fz : 1.0f - fz);
                atomicAdd(dst + slot, make_float2(wgt * gx, wgt * gy));
            }
        }
    }
}

// weight-norm fold + pack: one warp per output row, w = v * (g / ||v||) (torch._weight_norm).
struct PackLayer { const float* g; const float* v; const float* b; int rows, cols; int w_off, w_stride, w_transposed, b_off; };
struct PackArgs { PackLayer layer[5]; float* blob; };

__global__ void __launch_bounds__(256) pack_mlp_kernel(const PackArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int row = warp, li = 0;
    while (li < 5 && row >= a.layer[li].rows) { row -= a.layer[li].rows; ++li; }
    if (li >= 5) return;
    const PackLayer L = a.layer[li];
    float ss = 0.f;
    for (int c = lane; c < L.cols; c += 32) { const float v = L.v[row * L.cols + c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float sc = L.g[row] / sqrtf(ss);
    for (int c = lane; c < L.cols; c += 32) {
        const float w = L.v[row * L.cols + c] * sc;
        const int idx = L.w_transposed ? (L.w_off + c * L.w_stride + row) : (L.w_off + row * L.w_stride + c);
        a.blob[idx] = w;
    }
    if (L.b && lane == 0) a.blob[L.b_off + row] = L.b[row];
    // mirror the pieces the render kernel's epilogues read into the contiguous OFF_EPI block
    float* epi = a.blob + OFF_EPI;
    if (li == 0) {                                   // sdf layer 0: raw-xyz columns + bias
        if (lane < 3) epi[EPI_XB + 4 * row + lane] = L.v[row * L.cols + lane] * sc;
        if (lane == 3) epi[EPI_XB + 4 * row + 3] = L.b[row];
    } else if (li == 1) {                            // sdf layer 1, transposed, + bias
        for (int c = lane; c < L.cols; c += 32) epi[EPI_W1T + c * 16 + row] = L.v[row * L.cols + c] * sc;
        if (lane == 0) epi[EPI_B1 + row] = L.b[row];
    } else if (li == 4) {                            // colour head, transposed
        for (int c = lane; c < L.cols; c += 32) epi[EPI_C2T + c * 4 + row] = L.v[row * L.cols + c] * sc;
    }
}

__global__ void __launch_bounds__(256) upsample_round_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             const float* __restrict__ z, const float* __restrict__ sdf, uint32_t n,
                                                             uint32_t T, float inv_s, const float* __restrict__ alpha_in, float* alpha_out,
                                                             float* z_new_out, int32_t* bins, float* z_out, int32_t* order) {
    __shared__ float rows[kWarps * 4 * kMaxT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ray = blockIdx.x * kWarps + warp;
    if (ray >= n) return;
    float* zs = rows + warp * 4 * kMaxT; float* sdfs = zs + kMaxT; float* ta = sdfs + kMaxT; float* tb = ta + kMaxT;
    for (uint32_t k = lane; k < T; k += 32) { zs[k] = z[(size_t)ray * T + k]; sdfs[k] = sdf[(size_t)ray * T + k]; }
    __syncwarp();
    Ray r;
    r.ox = rays_o[3 * ray]; r.oy = rays_o[3 * ray + 1]; r.oz = rays_o[3 * ray + 2];
    r.dx = rays_d[3 * ray]; r.dy = rays_d[3 * ray + 1]; r.dz = rays_d[3 * ray + 2];
    float zn; int below, above;
    interval_alpha(r, zs, sdfs, ta, (int)T, inv_s, lane);
    if (alpha_out) for (uint32_t k = lane; k + 1 < T; k += 32) alpha_out[(size_t)ray * (T - 1) + k] = ta[k];
    if (alpha_in) for (uint32_t k = lane; k + 1 < T; k += 32) ta[k] = alpha_in[(size_t)ray * (T - 1) + k];
    __syncwarp();
    importance_from_alpha(zs, ta, tb, (int)T, lane, zn, below, above);
    int pos_old[4], pos_new;
    merge_positions(zs, (int)T, zn, lane, pos_old, pos_new);
    const size_t ob = (size_t)ray * (T + 16);
    if (lane < 16) {
        z_new_out[(size_t)ray * 16 + lane] = zn;
        bins[((size_t)ray * 16 + lane) * 2] = below; bins[((size_t)ray * 16 + lane) * 2 + 1] = above;
        z_out[ob + pos_new] = zn; order[ob + pos_new] = (int32_t)T + lane;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t k = lane + 32 * q;
        if (k < T) { z_out[ob + pos_old[q]] = zs[k]; order[ob + pos_old[q]] = (int32_t)k; }
    }
}

}  // namespace
namespace acb {
int launch_render_tc(const ac_nsr_model* m, const ac_nsr_render_args* a, cudaStream_t st);
int launch_forward_sdf_tc(const ac_nsr_model* m, const float* x, float* out, uint32_t B, float bound, cudaStream_t st, uint32_t stencil_M = 0,
                          float eps = 0.f, float* out_fd = nullptr, void* feat_cache = nullptr);
}
namespace {

int grid_for(uint32_t B, int block, int per_sm) {
    int sms = acb::sm_count();
    long want = ((long)B + block - 1) / block;
    long cap = (long)sms * per_sm;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

extern "C" {

int ac_nsr_pack_mlp(const float* sdf0_g, const float* sdf0_v, const float* sdf0_b, const float* sdf1_g,
                    const float* sdf1_v, const float* sdf1_b, const float* col0_g, const float* col0_v,
                    const float* col1_g, const float* col1_v, const float* col2_g, const float* col2_v,
                    float* blob, void* stream) {
    if (!sdf0_g || !sdf0_v || !sdf0_b || !sdf1_g || !sdf1_v || !sdf1_b || !col0_g || !col0_v || !col1_g ||
        !col1_v || !col2_g || !col2_v || !blob)
        return AC_E_INVALID_ARG;
    PackArgs a;
    a.layer[0] = {sdf0_g, sdf0_v, sdf0_b, 64, 35, OFF_W0, kSdfInPad, 0, OFF_B0};
    a.layer[1] = {sdf1_g, sdf1_v, sdf1_b, 16, 64, OFF_W1T, 16, 1, OFF_B1};
    a.layer[2] = {col0_g, col0_v, nullptr, 64, 21, OFF_C0, kColInPad, 0, 0};
    a.layer[3] = {col1_g, col1_v, nullptr, 64, 64, OFF_C1, kHidden, 0, 0};
    a.layer[4] = {col2_g, col2_v, nullptr, 3, 64, OFF_C2T, 4, 1, 0};
    a.blob = blob;
    // 211 weight rows -> 211 warps; padding columns are zeroed by the memset first.
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(blob, 0, BLOB_FLOATS * sizeof(float), st) != cudaSuccess) return acb::cuda_fail();
    pack_mlp_kernel<<<(211 * 32 + 255) / 256, 256, 0, st>>>(a);
    return acb::launched();
}

static int check_model(const ac_nsr_model* m) {
    if (!m || !m->embeddings || !m->offsets || !m->mlp_blob) return AC_E_INVALID_ARG;
    return AC_OK;
}

int ac_nsr_forward_sdf(const ac_nsr_model* m, const float* x, float* out, uint32_t B, float bound, void* stream) {
    if (check_model(m) || !x || !out) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    // Tensor-core kernel for anything but tiny batches (AC_SDF_IMPL=simt keeps the fp32 SIMT kernel for A/B debugging).
    static const bool use_simt = [] { const char* e = getenv("AC_SDF_IMPL"); return e && e[0] == 's'; }();
    if (!use_simt && B >= 4096) return acb::launch_forward_sdf_tc(m, x, out, B, bound, (cudaStream_t)stream);
    if (int rc = refresh_c_sdf(m->mlp_blob, (cudaStream_t)stream)) return rc;
    forward_sdf_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(m->embeddings), m->offsets, m->log2_per_level_scale,
        m->base_resolution, x, out, B, bound);
    return acb::launched();
}

int ac_nsr_forward_sdf_stencil(const ac_nsr_model* m, const float* P, uint32_t M, float bound, float eps, float* out_centre, float* out_fd,
                               void* stream) {
    if (check_model(m) || !P || !out_centre || !out_fd || !(eps > 0.f) || M > 0xFFFFFFFFu / 7u) return AC_E_INVALID_ARG;
    if (M == 0) return AC_OK;
    return acb::launch_forward_sdf_tc(m, P, out_centre, 7u * M, bound, (cudaStream_t)stream, M, eps, out_fd);
}

uint64_t ac_nsr_sdf_feature_cache_bytes(uint32_t n_points) { return (uint64_t)((n_points + 127u) / 128u) * 16384u; }

int ac_nsr_forward_sdf_stencil_cache(const ac_nsr_model* m, const float* P, uint32_t M, float bound, float eps, float* out_centre, float* out_fd,
                                     void* feature_cache, uint64_t cache_bytes, void* stream) {
    if (check_model(m) || !P || !out_centre || !out_fd || !(eps > 0.f) || M > 0xFFFFFFFFu / 7u) return AC_E_INVALID_ARG;
    if (!feature_cache || cache_bytes < ac_nsr_sdf_feature_cache_bytes(7u * M) || ((uintptr_t)feature_cache & 15)) return AC_E_WORKSPACE;
    if (M == 0) return AC_OK;
    return acb::launch_forward_sdf_tc(m, P, out_centre, 7u * M, bound, (cudaStream_t)stream, M, eps, out_fd, feature_cache);
}

int ac_nsr_sdf_backward(const ac_nsr_model* m, const float* x, const float* grad_out, uint32_t B, float bound,
                        float* grad_table, float* delta_a, float* hidden, float* feats, void* stream) {
    if (check_model(m) || !x || !grad_out || !grad_table || !delta_a || !hidden || !feats) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    constexpr size_t smem = kStageBytes + 64 * 32 * sizeof(float);
    ACB_SET_MAX_SMEM(sdf_backward_kernel, (int)smem);
    sdf_backward_kernel<<<grid_for(B, 256, 2), 256, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(m->embeddings), m->offsets, m->mlp_blob, m->log2_per_level_scale,
        m->base_resolution, x, grad_out, B, bound, grad_table, delta_a, hidden, feats);
    return acb::launched();
}

int ac_nsr_fd_gradient(const ac_nsr_model* m, const float* x, float* grad, uint32_t B, float bound, float epsilon,
                       void* stream) {
    if (check_model(m) || !x || !grad) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    ACB_SET_MAX_SMEM(fd_gradient_kernel, (int)kStageBytes);
    fd_gradient_kernel<<<grid_for(B, 256, 4), 256, kStageBytes, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(m->embeddings), m->offsets, m->mlp_blob, m->log2_per_level_scale,
        m->base_resolution, x, grad, B, bound, epsilon);
    return acb::launched();
}

int ac_nsr_forward_color(const ac_nsr_model* m, const float* x, const float* normal, const float* geo_feat, float* rgb,
                         uint32_t B, void* stream) {
    if (!m || !m->mlp_blob || !x || !normal || !geo_feat || !rgb) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    forward_color_kernel<<<(B + 255) / 256, 256, BLOB_FLOATS * sizeof(float), (cudaStream_t)stream>>>(
        m->mlp_blob, x, normal, geo_feat, rgb, B, nullptr);
    return acb::launched();
}

int ac_nsr_forward_color_bias(const ac_nsr_model* m, const float* x, const float* normal, const float* geo_feat, const float* c0_bias, float* rgb,
                              uint32_t B, void* stream) {
    if (!m || !m->mlp_blob || !x || !normal || !geo_feat || !c0_bias || !rgb) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    forward_color_kernel<<<(B + 255) / 256, 256, BLOB_FLOATS * sizeof(float), (cudaStream_t)stream>>>(
        m->mlp_blob, x, normal, geo_feat, rgb, B, c0_bias);
    return acb::launched();
}

int ac_nsr_viewdir_bias(const float* sh, const float* w_sh, uint32_t n, float* out, void* stream) {
    if (!sh || !w_sh || !out) return AC_E_INVALID_ARG;
    if (n == 0) return AC_OK;
    viewdir_bias_kernel<<<(unsigned)(((size_t)n * 64 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sh, w_sh, n, out);
    return acb::launched();
}

uint64_t ac_nsr_render_workspace_bytes(uint32_t n_rays) { return (uint64_t)n_rays * 2 * sizeof(float) + 16; }

int ac_nsr_render(const ac_nsr_model* m, const ac_nsr_render_args* a, void* stream) {
    if (check_model(m) || !m->variance || !a) return AC_E_INVALID_ARG;
    const bool sample_only = a->rgb == nullptr;          // documented in the header: z_vals is the only output
    if (!a->rays_o || !a->rays_d || !a->workspace) return AC_E_INVALID_ARG;
    if (sample_only ? (!a->z_vals || a->z_in) : (!a->depth || !a->weight_sum || !a->normal || !a->eikonal)) return AC_E_INVALID_ARG;
    const uint32_t T = a->num_steps + a->upsample_steps;
    if (a->num_steps < 2 || a->upsample_steps % 16 != 0 || T > (uint32_t)kMaxT) return AC_E_INVALID_ARG;
    if (a->workspace_bytes < ac_nsr_render_workspace_bytes(a->n_rays)) return AC_E_WORKSPACE;
    if (a->n_rays == 0) return AC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t seg = a->eikonal_segment ? a->eikonal_segment : a->n_rays;
    const char* impl_env = getenv("AC_RENDER_IMPL");
    const bool use_simt = impl_env && impl_env[0] == 's' && impl_env[1] == 'i';
    if (!use_simt) {   // tensor-core kernels (default); AC_RENDER_IMPL=simt keeps the SIMT kernel for A/B debugging
        int rc = acb::launch_render_tc(m, a, st);
        if (rc || sample_only) return rc;
        eikonal_reduce_kernel<<<(a->n_rays + seg - 1) / seg, 1024, 0, st>>>(reinterpret_cast<float*>(a->workspace), a->n_rays, seg, a->eikonal);
        return acb::launched();
    }
    if (a->z_in || a->pts_in || a->near_far_in || sample_only || a->c0_ray_bias) return AC_E_UNSUPPORTED;   // staged inputs / sampling-only / view directions: tensor-core kernel only
    RenderParams p;
    p.table = reinterpret_cast<const float2*>(m->embeddings);
    p.offsets = m->offsets; p.blob = m->mlp_blob; p.variance = m->variance;
    p.S = m->log2_per_level_scale; p.H = m->base_resolution;
    p.a = *a;
    p.eik_partial = reinterpret_cast<float*>(a->workspace);
    const size_t smem = kStageBytes + (size_t)kWarps * 4 * kMaxT * sizeof(float);
    ACB_SET_MAX_SMEM(nsr_render_kernel, (int)smem);
    nsr_render_kernel<<<(a->n_rays + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(p);
    int rc = acb::launched();
    if (rc) return rc;
    eikonal_reduce_kernel<<<(a->n_rays + seg - 1) / seg, 1024, 0, st>>>(p.eik_partial, a->n_rays, seg, a->eikonal);
    return acb::launched();
}

int ac_nsr_debug_upsample(const float* rays_o, const float* rays_d, const float* z, const float* sdf, uint32_t n_rays,
                          uint32_t T, float inv_s, const float* alpha_in, float* alpha_out, float* z_new, int32_t* bins,
                          float* z_out, int32_t* order, void* stream) {
    if (!rays_o || !rays_d || !z || !sdf || !z_new || !bins || !z_out || !order) return AC_E_INVALID_ARG;
    if (T < 2 || T + 16 > (uint32_t)kMaxT) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    upsample_round_kernel<<<(n_rays + kWarps - 1) / kWarps, kWarps * 32, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, z, sdf, n_rays, T, inv_s, alpha_in, alpha_out, z_new, bins, z_out, order);
    return acb::launched();
}

// One importance round as a stage-level operator (NeRFRenderer.up_sample + cat_z_vals, models/instant_nsr.py:410-475) for callers
// that evaluate the signed distance between rounds themselves -- the warped (render_can=False) path, whose coarse points go
// through the SMPL warp first (:166-172).  Same kernel as the parity-test entry above, computed section alphas.
int ac_nsr_upsample_round(const float* rays_o, const float* rays_d, const float* z, const float* sdf, uint32_t n_rays, uint32_t T, float inv_s,
                          float* z_new, int32_t* bins, float* z_out, int32_t* order, void* stream) {
    return ac_nsr_debug_upsample(rays_o, rays_d, z, sdf, n_rays, T, inv_s, nullptr, nullptr, z_new, bins, z_out, order, stream);
}

}  // extern "C"
