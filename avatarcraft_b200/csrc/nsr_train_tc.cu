// nsr_train_tc.cu -- fused backward of NeRFNetwork.forward_sdf with the weight gradients on the tensor cores (sm_100a).
//
// The first backward kernel (nsr_kernels.cu: sdf_backward_kernel) writes three per-point layer terms (delta_a [64,B],
// hidden [64,B], feats [36,B] = 2.4 GB per 4096-ray patch) that two cuBLAS GEMMs then reduce over the 3.7 M points.
// Here the reduction happens where the terms are produced:
//
//   per 128 points (one 4-warp group, thread = point):
//     A [128 x 128pts] fp16 = rows 0..63  delta_j * s_d        (dL/d hidden pre-activation, scaled into fp16 range)
//                             rows 64..127 hidden_j
//     B [ 64 x 128pts] fp16 = rows 0..35  layer input (x, y, z, 32 hash features, 1), rows 36..47 zero,
//                             rows 48..63 grad_out_o * s_g
//     D [128 x 64] (TMEM, fp32) += A B^T         8 x tcgen05.mma.kind::f16 M128 N64 K16, K = the 128 points
//   D[0..63][0..35]   = sum_pts delta_j in_i  = s_d [dW0 | db0]        (the ones row of the input gives the bias gradient)
//   D[64..127][48..63] = sum_pts hidden_j g_o = s_g dW1^T              (the other two quadrants are never read)
//
// The accumulator stays in TMEM for the whole persistent loop; each group adds its 3 328 useful entries to the global
// gradients once, at the end.  Operand precision: fp16 (11 significant bits, the class of the TF32 GEMMs this replaces),
// products exact in fp32, fp32 accumulation; the power-of-two scales s_d, s_g (device scalars computed by the caller from
// max|grad_out| and the layer-1 weights, no host sync) keep delta and grad_out inside the fp16 range whatever the loss
// scale is (the trainer multiplies one loss term by 1e5, stylize.py:190).
// Table scatter, forward recompute and arithmetic are those of sdf_backward_kernel.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"
#include "tc05.cuh"
#include "nsr_tc_group.cuh"

using namespace acb;

namespace {

constexpr int kGroupsT = 4;                              // 4-warp groups per CTA (512 threads, one CTA per SM)
constexpr int kThreadsT = 128 * kGroupsT;
constexpr int kMlpFloats = OFF_C0;                       // W0 [64][36] | B0 [64] | W1T [64][16] | B1 [16] -- all this kernel reads
// K-major no-swizzle tiles: 16-byte chunk (8 points) c of row r at c*LBO + r*16.  The chunk stride is programmable in the
// descriptor: 16 bytes of padding per plane (LBO = rows*16 + 16) put the four planes a warp's 32 points touch on distinct
// bank groups, so the per-point 2-byte column stores are conflict-free (they were 4-way conflicted at LBO = rows*16).
constexpr uint32_t A_LBO = 128 * 16 + 16, B_LBO = 64 * 16 + 16;
constexpr uint32_t A_BYTES = 16 * A_LBO, B_BYTES = 16 * B_LBO;
constexpr size_t T_SW = 0;
constexpr size_t T_LV = T_SW + kMlpFloats * sizeof(float);
constexpr size_t T_W0F = (T_LV + kLevels * sizeof(LevelMeta) + 15) / 16 * 16;
constexpr size_t T_TILES = (T_W0F + 64 * 32 * sizeof(float) + 1023) / 1024 * 1024;
constexpr size_t T_BARS = T_TILES + (size_t)kGroupsT * (A_BYTES + B_BYTES);
constexpr size_t T_TOTAL = T_BARS + kGroupsT * 8 + 16;
static_assert(T_TOTAL <= 227 * 1024, "shared memory budget");
static_assert(kMlpFloats % 4 == 0, "float4 staging");

// element (row r, point k) of a tile with chunk stride `lbo`
__device__ __forceinline__ void put(unsigned char* tile, uint32_t lbo, int r, int k, float v) {
    *reinterpret_cast<__half*>(tile + (k >> 3) * lbo + r * 16 + (k & 7) * 2) = __float2half_rn(v);
}

__global__ void __launch_bounds__(kThreadsT, 1) sdf_backward_tc_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                                       const float* __restrict__ blob, float S, uint32_t H,
                                                                       const float* __restrict__ x, const float* __restrict__ gout, uint32_t B,
                                                                       float bound, const float* __restrict__ scales, float* __restrict__ grad_table,
                                                                       float* __restrict__ grad_w0b, float* __restrict__ grad_w1, const uint32_t stencil_M,
                                                                       const float eps, const float* __restrict__ gout_fd) {
    // stencil_M > 0: x holds M section points, gout [M,16] their upstream gradients and gout_fd [6,M] the gradients of the
    // six finite-difference neighbours' signed distances; the B = 7 M points are generated here (see forward_sdf_tc_kernel).
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sw = reinterpret_cast<float*>(smem + T_SW);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + T_LV);
    float* w0f = reinterpret_cast<float*>(smem + T_W0F);          // [64][32] feature columns of W0, 16 B aligned rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroupsT);
    const int tid = threadIdx.x, warp = tid >> 5, group = warp >> 2, k = tid & 127;     // k = this thread's column (point) in its group's tiles
    unsigned char* At = smem + T_TILES + (size_t)group * (A_BYTES + B_BYTES);
    unsigned char* Bt = At + A_BYTES;

    for (int i = tid; i < kMlpFloats / 4; i += blockDim.x) reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(blob) + i);
    for (int i = tid; i < 64 * 32; i += blockDim.x) w0f[i] = __ldg(blob + OFF_W0 + (i >> 5) * kSdfInPad + 3 + (i & 31));
    if (tid < kLevels) lv[tid] = make_level_meta(offsets, tid, S, H, 3);
    if (tid == 0) {
        for (int g = 0; g < kGroupsT; ++g) tc05::mbar_init(bars + g, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<64 * kGroupsT>(tmem_slot);
    // rows 36..47 of B never change: zero them once (rows 0..35 and 48..63 are rewritten every round)
    for (int r = 36; r < 48; ++r) put(Bt, B_LBO, r, k, 0.f);
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tmem_slot + (uint32_t)group * 64u;      // this group's 64 accumulator columns
    const uint32_t a_s = tc05::smem_u32(At), b_s = tc05::smem_u32(Bt);
    const float s_d = scales[0], s_g = scales[1];
    uint32_t phase = 0;
    bool pending = false;

    for (uint32_t base = blockIdx.x * kThreadsT; base < B; base += gridDim.x * kThreadsT) {     // uniform trip count per CTA
        const uint32_t b = base + tid;
        const bool valid = b < B;
        float in[kSdfInPad];
        float g[16];
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) {
            uint32_t blk = 0, smp = b;
            if (stencil_M) { blk = b / stencil_M; smp = b - blk * stencil_M; }
            px = x[3 * (size_t)smp]; py = x[3 * (size_t)smp + 1]; pz = x[3 * (size_t)smp + 2];
            if (blk) {
                const float e = (blk & 1) ? eps : -eps;
                const uint32_t ax = (blk - 1) >> 1;
                if (ax == 0) px = clampf(px + e, -bound, bound);
                else if (ax == 1) py = clampf(py + e, -bound, bound);
                else pz = clampf(pz + e, -bound, bound);
            }
            encode_point(table, lv, bound, px, py, pz, in);
            if (blk == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(gout + 16 * (size_t)smp + 4 * q);
                    g[4 * q] = v.x; g[4 * q + 1] = v.y; g[4 * q + 2] = v.z; g[4 * q + 3] = v.w;
                }
            } else {
                g[0] = gout_fd[(size_t)(blk - 1) * stencil_M + smp];
#pragma unroll
                for (int q = 1; q < 16; ++q) g[q] = 0.f;
            }
        } else {
#pragma unroll
            for (int q = 0; q < kSdfInPad; ++q) in[q] = 0.f;
#pragma unroll
            for (int q = 0; q < 16; ++q) g[q] = 0.f;
        }
        if (pending) {                                   // the previous round's MMAs have consumed the tiles
            tc05::mbar_wait(bars + group, phase);
            phase ^= 1u;
            pending = false;
        }
        float din[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) din[q] = 0.f;
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
            const float h = softplus100_mufu(a);          // 2 MUFU ops, abs. error 1.7e-9 (as in the render kernel's hidden layer)
            const float4* __restrict__ w1 = reinterpret_cast<const float4*>(sw + OFF_W1T + j * 16);
            float dh = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w4 = w1[q];
                dh = fmaf(w4.x, g[4 * q], dh); dh = fmaf(w4.y, g[4 * q + 1], dh);
                dh = fmaf(w4.z, g[4 * q + 2], dh); dh = fmaf(w4.w, g[4 * q + 3], dh);
            }
            const float da = a * 100.0f > 20.0f ? dh : dh * sigmoid_mufu(a * 100.0f);
            put(At, A_LBO, j, k, valid ? da * s_d : 0.f);
            put(At, A_LBO, 64 + j, k, valid ? h : 0.f);
            const float4* __restrict__ wf = reinterpret_cast<const float4*>(w0f + j * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w4 = wf[q];
                din[4 * q + 0] = fmaf(w4.x, da, din[4 * q + 0]); din[4 * q + 1] = fmaf(w4.y, da, din[4 * q + 1]);
                din[4 * q + 2] = fmaf(w4.z, da, din[4 * q + 2]); din[4 * q + 3] = fmaf(w4.w, da, din[4 * q + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 35; ++q) put(Bt, B_LBO, q, k, in[q]);
        put(Bt, B_LBO, 35, k, valid ? 1.0f : 0.f);
#pragma unroll
        for (int q = 0; q < 16; ++q) put(Bt, B_LBO, 48 + q, k, g[q] * s_g);
        // hand the tiles to the tensor core: D += A B^T over this round's 128 points
        tc05::fence_proxy_async_smem();
        tc05::fence_before_sync();
        tc05::named_bar_sync(1 + group, 128);
        if (k == 0) {
            tc05::fence_after_sync();
            constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
#pragma unroll
            for (uint32_t s = 0; s < 8; ++s)             // K = 128 points = 8 x (K = 16): chunks 2s, 2s+1
                tc05::mma_f16(tmem, tc05::smem_desc(a_s + s * 2u * A_LBO, A_LBO, 128u), tc05::smem_desc(b_s + s * 2u * B_LBO, B_LBO, 128u),
                              idesc, (base != blockIdx.x * kThreadsT || s != 0) ? 1u : 0u);
            tc05::mma_commit(bars + group);
        }
        pending = true;
        if (!valid) continue;
        // scatter into the table (same arithmetic as kernel_grid_backward, hashencoder.cu:223-308)
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
            for (int c = 0; c < 8; ++c) {
                const uint32_t cx = ix + (c & 1), cy = iy + ((c >> 1) & 1), cz = iz + ((c >> 2) & 1);
                uint32_t slot;
                if (m.hashed == 0u) slot = cx + cy * m.res1 + cz * m.res1 * m.res1;
                else slot = wrap_slot(cx ^ (cy * 2654435761u) ^ (cz * 805459861u), m);
                const float wgt = (((c & 1) ? fx : 1.0f - fx) * ((c & 2) ? fy : 1.0f - fy)) * ((c & 4) ? fz : 1.0f - fz);
                atomicAdd(dst + slot, make_float2(wgt * gx, wgt * gy));
            }
        }
    }
    // ---- flush this group's accumulator: row = TMEM lane = unit ----
    if (pending) tc05::mbar_wait(bars + group, phase);
    tc05::fence_after_sync();
    if (blockIdx.x * kThreadsT < B) {                    // CTAs without work never issued an MMA: nothing to add
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int r = k;                                 // 0..127
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float acc[16];
            tc05::tmem_ld16(trow + q * 16, acc);
            if (r < 64) {                                // [dW0 | db0] * s_d: columns 0..35
                for (int c = 0; c < 16; ++c) { const int col = q * 16 + c; if (col < 36) atomicAdd(grad_w0b + r * 36 + col, acc[c]); }
            } else if (q == 3) {                         // dW1^T * s_g: columns 48..63 -> grad_w1[o][j]
                for (int c = 0; c < 16; ++c) atomicAdd(grad_w1 + c * 64 + (r - 64), acc[c]);
            }
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64 * kGroupsT>(*tmem_slot);
}


// ------------------------------------------------------------------------------------------------------------------------
// Second version (default): the per-point MLP recompute and back-propagation run on the tensor cores as well, with row =
// point tiles (the layout the render / forward kernels write), so the CUDA cores are left with the hash gathers, the
// activations and the table reductions.  Per 128 points (one 4-warp group, thread = point = tile row):
//
//   IN  [128 pts x 64]  fp16  cols 0..31 hash features (hi; lo in a second tile), 32..34 x y z, 35 = 1, 36..47 = 0,
//                             48..63 grad_out * s_g
//   round 1   a   [128 x 64] = IN_feat(hi,lo) W0f(hi,lo)^T                       6 x tcgen05.mma M128 N64 K16 (fp16x3)
//   epilogue  a += W0[:, xyz] p + b0 (fp32) ; h = softplus100(a) ; dh = W1^T grad_out (fp32: one product per unit for the six
//             finite-difference neighbours, whose 15 feature gradients are zero) ; delta = dh * sigmoid(100 a)
//   DH  [128 pts x 128] fp16  cols 0..63 delta * s_d (hi; lo in a second tile), 64..127 h
//   round 2   din [128 x 32] = DELTA(hi,lo) W0f(hi,lo)                           12 x tcgen05.mma M128 N32 K16 (fp16x3)
//             D   [128 x 64] += DH^T IN : the weight gradients, K = the 128 points  8 x tcgen05.mma M128 N64 K16
//   scatter   din / s_d into the eight corners of the sixteen levels (red.global.add.v2.f32)
//
// The weight-gradient product reads the SAME tiles transposed: a K-major [point][column] tile, seen as an MN-major operand,
// is [column][point] (core matrices of 8 points x 8 columns; SBO = the 2 KB column-chunk stride, LBO = the 128 B between
// 8-point groups), so nothing is written twice.  D[0..63][0..35] = s_d [dW0 (features, xyz) | db0], D[64..127][48..63] =
// s_g dW1^T, accumulated in TMEM over the persistent loop and flushed once per group.
// Three groups per CTA: each group needs 64 KB of tiles (IN 16, DH 32, lo 16) and 128 TMEM columns.
constexpr int kGroupsB = 3;
constexpr int kThreadsB = 128 * kGroupsB;
constexpr size_t M_XB = 0;                                        // float4 [64]: W0[j][x y z], b0[j]
constexpr size_t M_W1T = M_XB + 64 * 16;                          // float [64][16]
constexpr size_t M_LV = M_W1T + 64 * 16 * 4;
constexpr size_t M_WT = (M_LV + kLevels * sizeof(LevelMeta) + 1023) / 1024 * 1024;
constexpr uint32_t WT_W0_HI = 0, WT_W0_LO = 4096;                 // W0f  [n = 64 units][k = 32 features], chunk stride 1024
constexpr uint32_t WT_W0T_HI = 8192, WT_W0T_LO = 12288;           // W0f^T [n = 32 features][k = 64 units], chunk stride 512
constexpr uint32_t WT_BYTES = 16384;
constexpr size_t M_TILES = M_WT + WT_BYTES;
constexpr uint32_t G_IN = 0, G_DH = 16384, G_LO = 49152, G_BYTES = 65536;
constexpr size_t M_BARS = M_TILES + (size_t)kGroupsB * G_BYTES;
constexpr size_t M_TOTAL = M_BARS + kGroupsB * 8 + 16;
static_assert(M_TOTAL <= 227 * 1024, "shared memory budget");

// kind::f16, fp32 accumulate, A and B both MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_f16_mn(uint32_t M, uint32_t N) { return tc05::idesc_f16(M, N) | (1u << 15) | (1u << 16); }

// SPLIT = true: the table scatter is left to sdf_scatter_kernel; this kernel writes dL/d(features) of every point to
// din_t [32][B] (feature-major, so both kernels' accesses are coalesced) instead.
template <bool SPLIT>
__global__ void __launch_bounds__(kThreadsB, 1) sdf_backward_mma_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                                        const float* __restrict__ blob, float S, uint32_t H,
                                                                        const float* __restrict__ x, const float* __restrict__ gout, uint32_t B,
                                                                        float bound, const float* __restrict__ scales, float* __restrict__ grad_table,
                                                                        float* __restrict__ grad_w0b, float* __restrict__ grad_w1, const uint32_t stencil_M,
                                                                        const float eps, const float* __restrict__ gout_fd, float* __restrict__ din_t,
                                                                        const unsigned char* __restrict__ feat_cache) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float4* xb = reinterpret_cast<float4*>(smem + M_XB);
    float* w1t = reinterpret_cast<float*>(smem + M_W1T);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + M_LV);
    unsigned char* wt = smem + M_WT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroupsB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, group = warp >> 2;

    for (int j = tid; j < 64; j += blockDim.x)
        xb[j] = make_float4(__ldg(blob + OFF_W0 + j * kSdfInPad), __ldg(blob + OFF_W0 + j * kSdfInPad + 1), __ldg(blob + OFF_W0 + j * kSdfInPad + 2),
                            __ldg(blob + OFF_B0 + j));
    for (int i = tid; i < 64 * 16; i += blockDim.x) w1t[i] = __ldg(blob + OFF_W1T + i);
    if (tid < kLevels) lv[tid] = make_level_meta(offsets, tid, S, H, 3);
    for (int i = tid; i < 64 * 32; i += blockDim.x) {
        const int j = i >> 5, f = i & 31;
        const float w = __ldg(blob + OFF_W0 + j * kSdfInPad + 3 + f);
        stage_b_tile(wt + WT_W0_HI, wt + WT_W0_LO, j, f, w);
        const __half h = __float2half_rn(w), l = __float2half_rn(w - __half2float(h));
        const int at = (j >> 3) * 512 + f * 16 + (j & 7) * 2;
        *reinterpret_cast<__half*>(wt + WT_W0T_HI + at) = h;
        *reinterpret_cast<__half*>(wt + WT_W0T_LO + at) = l;
    }
    if (tid == 0) {
        for (int g = 0; g < kGroupsB; ++g) tc05::mbar_init(bars + g, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<512>(tmem_slot);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    Group g;
    g.a = smem + M_TILES + (size_t)group * G_BYTES;
    g.a_s = tc05::smem_u32(g.a);
    g.b_s = tc05::smem_u32(wt);
    g.bar = bars + group;
    g.phase = 0;
    g.row = (warp & 3) * 32 + lane;
    g.tmem = tmem_base + (uint32_t)group * 128u + ((uint32_t)((warp & 3) * 32) << 16);
    g.bar_id = 1 + group;
    {
        bool ok = true;
        for (int l = 0; l < kLevels; ++l) ok = ok && (lv[l].hashed == (l < 5 ? 0u : 1u));
        g.std_layout = ok;
    }
    const uint32_t tmem_acc = (g.tmem & 0xFFFFu) + 64u;                  // weight-gradient accumulator columns of this group
    const float s_d = scales[0], s_g = scales[1], inv_s_d = 1.0f / s_d;
    unsigned char* row_ptr = g.a + g.row * 16;

    for (uint32_t base = blockIdx.x * kThreadsB; base < B; base += gridDim.x * kThreadsB) {     // uniform trip count per CTA
        const uint32_t b = base + tid;
        const bool valid = b < B;
        float px = 3.0f * bound + 1.0f, py = px, pz = px;                 // padding threads: out of range -> zero features, no scatter
        uint32_t blk = 0, smp = 0;
        float gq[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) gq[q] = 0.f;
        if (valid) {
            smp = b;
            if (stencil_M) { blk = b / stencil_M; smp = b - blk * stencil_M; }
            px = x[3 * (size_t)smp]; py = x[3 * (size_t)smp + 1]; pz = x[3 * (size_t)smp + 2];
            if (blk) {
                const float e = (blk & 1) ? eps : -eps;
                const uint32_t ax = (blk - 1) >> 1;
                if (ax == 0) px = clampf(px + e, -bound, bound);
                else if (ax == 1) py = clampf(py + e, -bound, bound);
                else pz = clampf(pz + e, -bound, bound);
                gq[0] = gout_fd[(size_t)(blk - 1) * stencil_M + smp];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(gout + 16 * (size_t)smp + 4 * q);
                    gq[4 * q] = v.x; gq[4 * q + 1] = v.y; gq[4 * q + 2] = v.z; gq[4 * q + 3] = v.w;
                }
            }
        }
        const bool sd_only = __all_sync(0xffffffffu, blk != 0u || !valid);   // warp-uniform: only d/d(signed distance) is non-zero
        // ---- IN tile: features (hi | lo), then (x y z 1 0 0 0 0), zeros, grad_out * s_g ----
        if (feat_cache) {        // the forward of the same points left this tile's encoded features (hi | lo chunks): no gathers
            const uint32_t g_first = base + (uint32_t)group * 128u;         // a trailing group past the last point has no tile
            const unsigned char* src = feat_cache + (size_t)(g_first >> 7) * 16384 + g.row * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(row_ptr + G_IN + c * 2048) = g_first < B ? __ldg(reinterpret_cast<const uint4*>(src + c * 2048)) : z4;
                *reinterpret_cast<uint4*>(row_ptr + G_LO + c * 2048) = g_first < B ? __ldg(reinterpret_cast<const uint4*>(src + 8192 + c * 2048)) : z4;
            }
        } else {
            encode_to_tile<G_LO>(row_ptr + G_IN, table, lv, bound, px, py, pz, g.std_layout);
        }
        {
            uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
            if (valid) { c4.x = tc05::pack_f16x2(px, py); c4.y = tc05::pack_f16x2(pz, 1.0f); }
            *reinterpret_cast<uint4*>(row_ptr + G_IN + 4 * 2048) = c4;
            *reinterpret_cast<uint4*>(row_ptr + G_IN + 5 * 2048) = make_uint4(0u, 0u, 0u, 0u);
            uint4 g0, g1;
            g0.x = tc05::pack_f16x2(gq[0] * s_g, gq[1] * s_g); g0.y = tc05::pack_f16x2(gq[2] * s_g, gq[3] * s_g);
            g0.z = tc05::pack_f16x2(gq[4] * s_g, gq[5] * s_g); g0.w = tc05::pack_f16x2(gq[6] * s_g, gq[7] * s_g);
            g1.x = tc05::pack_f16x2(gq[8] * s_g, gq[9] * s_g); g1.y = tc05::pack_f16x2(gq[10] * s_g, gq[11] * s_g);
            g1.z = tc05::pack_f16x2(gq[12] * s_g, gq[13] * s_g); g1.w = tc05::pack_f16x2(gq[14] * s_g, gq[15] * s_g);
            *reinterpret_cast<uint4*>(row_ptr + G_IN + 6 * 2048) = g0;
            *reinterpret_cast<uint4*>(row_ptr + G_IN + 7 * 2048) = g1;
        }
        // ---- round 1: pre-activation of the hidden layer (feature part) ----
        group_mma_round(g, [&] {
            constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
#pragma unroll
            for (uint32_t s = 0; s < 2; ++s) {
                const uint64_t ah = tc05::smem_desc(g.a_s + G_IN + s * 4096u, 2048u, 128u);
                const uint64_t al = tc05::smem_desc(g.a_s + G_LO + s * 4096u, 2048u, 128u);
                const uint64_t bh = tc05::smem_desc(g.b_s + WT_W0_HI + s * 2048u, 1024u, 128u);
                const uint64_t bl = tc05::smem_desc(g.b_s + WT_W0_LO + s * 2048u, 1024u, 128u);
                tc05::mma_f16(g.tmem & 0xFFFFu, ah, bh, idesc, s);
                tc05::mma_f16(g.tmem & 0xFFFFu, al, bh, idesc, 1u);
                tc05::mma_f16(g.tmem & 0xFFFFu, ah, bl, idesc, 1u);
            }
        });
        // ---- epilogue: activation, its derivative, dL/d(pre-activation) -> DH tile ----
#pragma unroll 1
        for (int qtr = 0; qtr < 4; ++qtr) {
            float acc[16];
            tc05::tmem_ld16(g.tmem + qtr * 16, acc);
            float dl[16], hh[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int j = qtr * 16 + jj;
                const float4 w = xb[j];
                const float a = acc[jj] + fmaf(w.x, px, fmaf(w.y, py, fmaf(w.z, pz, w.w)));
                hh[jj] = softplus100_mufu(a);
                float dh;
                if (sd_only) {
                    dh = w1t[j * 16] * gq[0];
                } else {
                    const float4* __restrict__ w1 = reinterpret_cast<const float4*>(w1t + j * 16);
                    dh = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w4 = w1[q];
                        dh = fmaf(w4.x, gq[4 * q], dh); dh = fmaf(w4.y, gq[4 * q + 1], dh);
                        dh = fmaf(w4.z, gq[4 * q + 2], dh); dh = fmaf(w4.w, gq[4 * q + 3], dh);
                    }
                }
                const float da = a * 100.0f > 20.0f ? dh : dh * sigmoid_mufu(a * 100.0f);
                dl[jj] = da * s_d;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint4 hi, lo, hd;
                tc05::split_f16x2(dl[8 * half + 0], dl[8 * half + 1], hi.x, lo.x); tc05::split_f16x2(dl[8 * half + 2], dl[8 * half + 3], hi.y, lo.y);
                tc05::split_f16x2(dl[8 * half + 4], dl[8 * half + 5], hi.z, lo.z); tc05::split_f16x2(dl[8 * half + 6], dl[8 * half + 7], hi.w, lo.w);
                hd.x = tc05::pack_f16x2(hh[8 * half + 0], hh[8 * half + 1]); hd.y = tc05::pack_f16x2(hh[8 * half + 2], hh[8 * half + 3]);
                hd.z = tc05::pack_f16x2(hh[8 * half + 4], hh[8 * half + 5]); hd.w = tc05::pack_f16x2(hh[8 * half + 6], hh[8 * half + 7]);
                const int c = 2 * qtr + half;
                *reinterpret_cast<uint4*>(row_ptr + G_DH + c * 2048) = hi;
                *reinterpret_cast<uint4*>(row_ptr + G_LO + c * 2048) = lo;
                *reinterpret_cast<uint4*>(row_ptr + G_DH + (8 + c) * 2048) = hd;
            }
        }
        // ---- round 2: dL/d(features) and the weight-gradient accumulation ----
        const bool first = base == blockIdx.x * kThreadsB;
        group_mma_round(g, [&] {
            constexpr uint32_t idesc_d = tc05::idesc_f16(128, 32);
#pragma unroll
            for (uint32_t s = 0; s < 4; ++s) {
                const uint64_t ah = tc05::smem_desc(g.a_s + G_DH + s * 4096u, 2048u, 128u);
                const uint64_t al = tc05::smem_desc(g.a_s + G_LO + s * 4096u, 2048u, 128u);
                const uint64_t bh = tc05::smem_desc(g.b_s + WT_W0T_HI + s * 1024u, 512u, 128u);
                const uint64_t bl = tc05::smem_desc(g.b_s + WT_W0T_LO + s * 1024u, 512u, 128u);
                tc05::mma_f16(g.tmem & 0xFFFFu, ah, bh, idesc_d, s);
                tc05::mma_f16(g.tmem & 0xFFFFu, al, bh, idesc_d, 1u);
                tc05::mma_f16(g.tmem & 0xFFFFu, ah, bl, idesc_d, 1u);
            }
            constexpr uint32_t idesc_w = idesc_f16_mn(128, 64);
#pragma unroll
            for (uint32_t s = 0; s < 8; ++s)             // K = the 128 points = 8 x (K = 16): 8-point groups 2s, 2s+1
                tc05::mma_f16(tmem_acc, tc05::smem_desc(g.a_s + G_DH + s * 256u, 128u, 2048u), tc05::smem_desc(g.a_s + G_IN + s * 256u, 128u, 2048u),
                              idesc_w, (first && s == 0) ? 0u : 1u);
        });
        float din[32];
        {
            float acc[16];
            tc05::tmem_ld16(g.tmem, acc);
#pragma unroll
            for (int q = 0; q < 16; ++q) din[q] = acc[q] * inv_s_d;
            tc05::tmem_ld16(g.tmem + 16, acc);
#pragma unroll
            for (int q = 0; q < 16; ++q) din[16 + q] = acc[q] * inv_s_d;
        }
        if constexpr (SPLIT) {
            if (valid) {
#pragma unroll
                for (int q = 0; q < 32; ++q) din_t[(size_t)q * B + b] = din[q];
            }
            continue;
        }
        // ---- scatter into the table (same arithmetic as kernel_grid_backward, hashencoder.cu:223-308) ----
        const float two_b = 2.0f * bound;
        const float u = (px + bound) / two_b, v = (py + bound) / two_b, w = (pz + bound) / two_b;
        if (!valid || (u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f)) continue;
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
            for (int c = 0; c < 8; ++c) {
                const uint32_t cx = ix + (c & 1), cy = iy + ((c >> 1) & 1), cz = iz + ((c >> 2) & 1);
                uint32_t slot;
                if (m.hashed == 0u) slot = cx + cy * m.res1 + cz * m.res1 * m.res1;
                else slot = wrap_slot(cx ^ (cy * 2654435761u) ^ (cz * 805459861u), m);
                const float wgt = (((c & 1) ? fx : 1.0f - fx) * ((c & 2) ? fy : 1.0f - fy)) * ((c & 4) ? fz : 1.0f - fz);
                atomicAdd(dst + slot, make_float2(wgt * gx, wgt * gy));
            }
        }
    }
    // ---- flush this group's accumulator: row = TMEM lane = unit; columns in IN-tile order ----
    tc05::fence_after_sync();
    if (blockIdx.x * kThreadsB < B) {                    // CTAs without work never issued an MMA: nothing to add
        const int r = g.row;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            float acc[16];
            tc05::tmem_ld16(g.tmem + 64 + q * 16, acc);
            if (r < 64) {                                // s_d [dW0 | db0]: features -> columns 3..34, xyz -> 0..2, the ones column -> 35
                for (int c = 0; c < 16; ++c) {
                    const int col = q * 16 + c;
                    if (col < 32) atomicAdd(grad_w0b + r * 36 + 3 + col, acc[c]);
                    else if (col < 35) atomicAdd(grad_w0b + r * 36 + (col - 32), acc[c]);
                    else if (col == 35) atomicAdd(grad_w0b + r * 36 + 35, acc[c]);
                }
            } else if (q == 3) {                         // s_g dW1^T: columns 48..63 -> grad_w1[o][j]
                for (int c = 0; c < 16; ++c) atomicAdd(grad_w1 + c * 64 + (r - 64), acc[c]);
            }
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

// The two x-corners of a cell are neighbouring table entries more often than not (slot, slot + 1 on a dense level; s, s ^ 1 on
// a hashed level whenever the cell's x index is even): when the pair also starts on a 16-byte boundary, ONE
// red.global.add.v4.f32 carries both corners' contributions -- one packet on the SM's reduction path instead of two.  (With the
// reference's offsets table the hashed-level pairs are aligned iff the gradient table starts 8 bytes past a 16-byte boundary,
// which is where utils/optim.FlatAdam places it; any other placement just takes the two-packet path more often.)
__device__ __forceinline__ void red_corner_pair(float2* __restrict__ dst, uint32_t slot0, uint32_t slot1, float ax, float ay, float bx, float by) {
    const uint32_t lo = slot0 < slot1 ? slot0 : slot1;
    float2* p = dst + lo;
    if ((slot0 ^ slot1) == 1u && (reinterpret_cast<uintptr_t>(p) & 15u) == 0u) {
        const bool first = lo == slot0;
        atomicAdd(reinterpret_cast<float4*>(p), first ? make_float4(ax, ay, bx, by) : make_float4(bx, by, ax, ay));
    } else {
        atomicAdd(dst + slot0, make_float2(ax, ay));
        atomicAdd(dst + slot1, make_float2(bx, by));
    }
}

// The table scatter of the split backward: grad_table[corner] += w_corner * din (hashencoder.cu:223-308) for every point and
// level.  A block takes 256 consecutive points; warp w handles levels 2w and 2w+1 of all of them, 32 consecutive points per
// instruction (consecutive samples of a ray: the din_t reads are coalesced).  No shared-memory tiles, full occupancy: this
// kernel is bound by the SM's reduction path to L2 alone, the gather / MMA kernel no longer waits behind it.
__global__ void __launch_bounds__(256) sdf_scatter_kernel(const int32_t* __restrict__ offsets, float S, uint32_t H, const float* __restrict__ x,
                                                          uint32_t B, float bound, uint32_t stencil_M, float eps, const float* __restrict__ din_t,
                                                          float* __restrict__ grad_table) {
    __shared__ LevelMeta lv[kLevels];
    if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(offsets, threadIdx.x, S, H, 3);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const float two_b = 2.0f * bound;
    for (uint32_t base = blockIdx.x * 256u; base < B; base += gridDim.x * 256u) {
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const uint32_t b = base + it * 32 + lane;            // warp-uniform trip counts: every lane joins the shuffles below
            bool inr = b < B;
            float u = 0.f, v = 0.f, w = 0.f;
            if (inr) {
                uint32_t blk = 0, smp = b;
                if (stencil_M) { blk = b / stencil_M; smp = b - blk * stencil_M; }
                float px = x[3 * (size_t)smp], py = x[3 * (size_t)smp + 1], pz = x[3 * (size_t)smp + 2];
                if (blk) {
                    const float e = (blk & 1) ? eps : -eps;
                    const uint32_t ax = (blk - 1) >> 1;
                    if (ax == 0) px = clampf(px + e, -bound, bound);
                    else if (ax == 1) py = clampf(py + e, -bound, bound);
                    else pz = clampf(pz + e, -bound, bound);
                }
                u = (px + bound) / two_b; v = (py + bound) / two_b; w = (pz + bound) / two_b;
                inr = !((u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f));
            }
            if (!__any_sync(full, inr)) continue;
#pragma unroll 1
            for (int ll = 0; ll < 2; ++ll) {
                const int l = 2 * warp + ll;
                const LevelMeta m = lv[l];
                float gx = 0.f, gy = 0.f;
                if (inr) { gx = din_t[(size_t)(2 * l) * B + b]; gy = din_t[(size_t)(2 * l + 1) * B + b]; }
                float fx = fmaf(u, m.scale, 0.5f), fy = fmaf(v, m.scale, 0.5f), fz = fmaf(w, m.scale, 0.5f);
                const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
                const uint32_t ix = (uint32_t)flx, iy = (uint32_t)fly, iz = (uint32_t)flz;
                fx -= flx; fy -= fly; fz -= flz;
                float val[16];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float wgt = (((c & 1) ? fx : 1.0f - fx) * ((c & 2) ? fy : 1.0f - fy)) * ((c & 4) ? fz : 1.0f - fz);
                    val[2 * c] = wgt * gx; val[2 * c + 1] = wgt * gy;
                }
                // Lanes are consecutive samples of one ray: those inside the same grid cell are a contiguous run and add into
                // the same eight corners.  Sum each run in registers (segmented scan) and let its last lane issue the
                // reductions: one packet per (run, corner) instead of one per (sample, corner).
                const uint32_t k1 = inr ? (ix | (iy << 16)) : 0xFFFFFFFFu, k2 = inr ? iz : (0x80000000u | (uint32_t)lane);
                const uint32_t p1 = __shfl_up_sync(full, k1, 1), p2 = __shfl_up_sync(full, k2, 1);
                const bool head = lane == 0 || k1 != p1 || k2 != p2;
                const unsigned heads = __ballot_sync(full, head);
                bool emit = inr;
                if (heads != full) {
                    const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
                    emit = inr && (lane == 31 || ((heads >> (lane + 1)) & 1u));
                    // shuffles share the L1 data pipe with the reductions (ncu: half of its wavefronts): run only the scan
                    // steps the warp's longest run needs -- one or two at the fine levels instead of five
                    const int longest = (int)__reduce_max_sync(full, (unsigned)(lane - start));
#pragma unroll 1
                    for (int d = 1; d <= longest; d <<= 1) {
                        const bool take = lane - d >= start;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float t = __shfl_up_sync(full, val[q], d);
                            if (take) val[q] += t;
                        }
                    }
                }
                if (!emit) continue;
                float2* __restrict__ dst = reinterpret_cast<float2*>(grad_table) + m.offset;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {                 // corner pairs (x, x + 1) at fixed (y, z)
                    const uint32_t cy = iy + ((c >> 1) & 1), cz = iz + ((c >> 2) & 1);
                    uint32_t s0, s1;
                    if (m.hashed == 0u) { s0 = ix + cy * m.res1 + cz * m.res1 * m.res1; s1 = s0 + 1u; }
                    else {
                        const uint32_t hyz = (cy * 2654435761u) ^ (cz * 805459861u);
                        s0 = wrap_slot(ix ^ hyz, m); s1 = wrap_slot((ix + 1u) ^ hyz, m);
                    }
                    red_corner_pair(dst, s0, s1, val[2 * c], val[2 * c + 1], val[2 * c + 2], val[2 * c + 3]);
                }
            }
        }
    }
}

// Stencil variant of the scatter: the seven points of a sample's finite-difference stencil (centre, +-0.005 along x, y, z) lie in
// one grid cell on the coarse levels, and so do the stencils of the next samples of the ray.  Lanes are laid out as
// (sample, stencil point) -- 4 consecutive samples x 8 lanes (7 points + a zero-weight copy of the centre) -- so a run of
// equal cells now spans whole stencils: one reduction packet per (run, corner) where the point-ordered kernel above
// needs one per stencil block.  Same arithmetic, different summation order.
__global__ void __launch_bounds__(256) sdf_scatter_stencil_kernel(const int32_t* __restrict__ offsets, float S, uint32_t H, const float* __restrict__ x,
                                                                  uint32_t M, float bound, float eps, const float* __restrict__ din_t,
                                                                  float* __restrict__ grad_table) {
    __shared__ LevelMeta lv[kLevels];
    if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(offsets, threadIdx.x, S, H, 3);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const float two_b = 2.0f * bound;
    const size_t B = (size_t)7 * M;
    const uint32_t sub = (uint32_t)lane >> 3, st = (uint32_t)lane & 7u;
    const uint32_t blk = st == 7u ? 0u : st;                     // lane 7 of a sample: a copy of the centre that carries no gradient
    for (uint32_t base = blockIdx.x * 32u; base < M; base += gridDim.x * 32u) {
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const uint32_t smp = base + it * 4 + sub;
            bool inr = smp < M;
            float u = 0.f, v = 0.f, w = 0.f;
            if (inr) {
                float px = x[3 * (size_t)smp], py = x[3 * (size_t)smp + 1], pz = x[3 * (size_t)smp + 2];
                if (blk) {
                    const float e = (blk & 1) ? eps : -eps;
                    const uint32_t ax = (blk - 1) >> 1;
                    if (ax == 0) px = clampf(px + e, -bound, bound);
                    else if (ax == 1) py = clampf(py + e, -bound, bound);
                    else pz = clampf(pz + e, -bound, bound);
                }
                u = (px + bound) / two_b; v = (py + bound) / two_b; w = (pz + bound) / two_b;
                inr = !((u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f));
            }
            if (!__any_sync(full, inr)) continue;
            const size_t b = (size_t)blk * M + smp;
#pragma unroll 1
            for (int ll = 0; ll < 2; ++ll) {
                const int l = 2 * warp + ll;
                const LevelMeta m = lv[l];
                float gx = 0.f, gy = 0.f;
                if (inr && st != 7u) { gx = din_t[(size_t)(2 * l) * B + b]; gy = din_t[(size_t)(2 * l + 1) * B + b]; }
                float fx = fmaf(u, m.scale, 0.5f), fy = fmaf(v, m.scale, 0.5f), fz = fmaf(w, m.scale, 0.5f);
                const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
                const uint32_t ix = (uint32_t)flx, iy = (uint32_t)fly, iz = (uint32_t)flz;
                fx -= flx; fy -= fly; fz -= flz;
                float val[16];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float wgt = (((c & 1) ? fx : 1.0f - fx) * ((c & 2) ? fy : 1.0f - fy)) * ((c & 4) ? fz : 1.0f - fz);
                    val[2 * c] = wgt * gx; val[2 * c + 1] = wgt * gy;
                }
                const uint32_t k1 = inr ? (ix | (iy << 16)) : 0xFFFFFFFFu, k2 = inr ? iz : (0x80000000u | (uint32_t)lane);
                const uint32_t p1 = __shfl_up_sync(full, k1, 1), p2 = __shfl_up_sync(full, k2, 1);
                const bool head = lane == 0 || k1 != p1 || k2 != p2;
                const unsigned heads = __ballot_sync(full, head);
                bool emit = inr;
                if (heads != full) {
                    const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
                    emit = inr && (lane == 31 || ((heads >> (lane + 1)) & 1u));
                    // shuffles share the L1 data pipe with the reductions (ncu: half of its wavefronts): run only the scan
                    // steps the warp's longest run needs -- one or two at the fine levels instead of five
                    const int longest = (int)__reduce_max_sync(full, (unsigned)(lane - start));
#pragma unroll 1
                    for (int d = 1; d <= longest; d <<= 1) {
                        const bool take = lane - d >= start;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const float t = __shfl_up_sync(full, val[q], d);
                            if (take) val[q] += t;
                        }
                    }
                    if (st == 7u && head) emit = false;          // a zero-weight copy that is a run of its own has nothing to add
                } else if (st == 7u) {
                    emit = false;
                }
                if (!emit) continue;
                float2* __restrict__ dst = reinterpret_cast<float2*>(grad_table) + m.offset;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {                 // corner pairs (x, x + 1) at fixed (y, z)
                    const uint32_t cy = iy + ((c >> 1) & 1), cz = iz + ((c >> 2) & 1);
                    uint32_t s0, s1;
                    if (m.hashed == 0u) { s0 = ix + cy * m.res1 + cz * m.res1 * m.res1; s1 = s0 + 1u; }
                    else {
                        const uint32_t hyz = (cy * 2654435761u) ^ (cz * 805459861u);
                        s0 = wrap_slot(ix ^ hyz, m); s1 = wrap_slot((ix + 1u) ^ hyz, m);
                    }
                    red_corner_pair(dst, s0, s1, val[2 * c], val[2 * c + 1], val[2 * c + 2], val[2 * c + 3]);
                }
            }
        }
    }
}

bool use_v1_backward() {
    static const bool v1 = [] { const char* e = getenv("AC_SDF_BWD_IMPL"); return e && e[0] == 'v' && e[1] == '1'; }();
    return v1;
}

int launch_sdf_backward(const ac_nsr_model* m, const float* x, const float* gout, uint32_t B, float bound, const float* scales, float* grad_table,
                        float* grad_w0b, float* grad_w1, uint32_t stencil_M, float eps, const float* gout_fd, cudaStream_t st,
                        void* workspace = nullptr, uint64_t workspace_bytes = 0, const void* feature_cache = nullptr) {
    const float2* table = reinterpret_cast<const float2*>(m->embeddings);
    if (use_v1_backward()) {
        ACB_SET_MAX_SMEM(sdf_backward_tc_kernel, T_TOTAL);
        const uint32_t want = (B + kThreadsT - 1) / kThreadsT;
        const uint32_t grid = want < (uint32_t)acb::sm_count() ? want : (uint32_t)acb::sm_count();
        sdf_backward_tc_kernel<<<grid, kThreadsT, T_TOTAL, st>>>(table, m->offsets, m->mlp_blob, m->log2_per_level_scale, m->base_resolution, x, gout, B,
                                                                 bound, scales, grad_table, grad_w0b, grad_w1, stencil_M, eps, gout_fd);
    } else {
        const uint32_t want = (B + kThreadsB - 1) / kThreadsB;
        const uint32_t grid = want < (uint32_t)acb::sm_count() ? want : (uint32_t)acb::sm_count();
        if (workspace && workspace_bytes >= (uint64_t)B * 32u * sizeof(float)) {
            // split: gather + MMA kernel writes dL/d(features) to the workspace, a full-occupancy kernel scatters it
            float* din_t = reinterpret_cast<float*>(workspace);
            ACB_SET_MAX_SMEM(sdf_backward_mma_kernel<true>, M_TOTAL);
            sdf_backward_mma_kernel<true><<<grid, kThreadsB, M_TOTAL, st>>>(table, m->offsets, m->mlp_blob, m->log2_per_level_scale, m->base_resolution, x,
                                                                            gout, B, bound, scales, grad_table, grad_w0b, grad_w1, stencil_M, eps, gout_fd,
                                                                            din_t, reinterpret_cast<const unsigned char*>(feature_cache));
            if (int rc = acb::launched()) return rc;
            const uint32_t cap = (uint32_t)acb::sm_count() * 8u;
            static const bool point_order = [] { const char* e = getenv("AC_SCATTER_ORDER"); return e && e[0] == 'p'; }();   // A/B
            if (stencil_M && !point_order) {
                const uint32_t blocks = (stencil_M + 31u) / 32u;
                sdf_scatter_stencil_kernel<<<blocks < cap ? blocks : cap, 256, 0, st>>>(m->offsets, m->log2_per_level_scale, m->base_resolution, x,
                                                                                       stencil_M, bound, eps, din_t, grad_table);
            } else {
                const uint32_t blocks = (B + 255u) / 256u;
                sdf_scatter_kernel<<<blocks < cap ? blocks : cap, 256, 0, st>>>(m->offsets, m->log2_per_level_scale, m->base_resolution, x, B, bound,
                                                                               stencil_M, eps, din_t, grad_table);
            }
            return acb::launched();
        }
        ACB_SET_MAX_SMEM(sdf_backward_mma_kernel<false>, M_TOTAL);
        sdf_backward_mma_kernel<false><<<grid, kThreadsB, M_TOTAL, st>>>(table, m->offsets, m->mlp_blob, m->log2_per_level_scale, m->base_resolution, x,
                                                                         gout, B, bound, scales, grad_table, grad_w0b, grad_w1, stencil_M, eps, gout_fd,
                                                                         nullptr, reinterpret_cast<const unsigned char*>(feature_cache));
    }
    return acb::launched();
}

}  // namespace

extern "C" int ac_nsr_sdf_backward_fused(const ac_nsr_model* m, const float* x, const float* grad_out, uint32_t B, float bound,
                                         const float* scales, float* grad_table, float* grad_w0b, float* grad_w1, void* stream) {
    if (!m || !m->embeddings || !m->offsets || !m->mlp_blob || !x || !grad_out || !scales || !grad_table || !grad_w0b || !grad_w1)
        return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    return launch_sdf_backward(m, x, grad_out, B, bound, scales, grad_table, grad_w0b, grad_w1, 0u, 0.f, nullptr, (cudaStream_t)stream);
}

extern "C" int ac_nsr_sdf_backward_stencil(const ac_nsr_model* m, const float* P, uint32_t M, float bound, float eps, const float* grad_centre,
                                           const float* grad_fd, const float* scales, float* grad_table, float* grad_w0b, float* grad_w1,
                                           void* stream) {
    if (!m || !m->embeddings || !m->offsets || !m->mlp_blob || !P || !grad_centre || !grad_fd || !scales || !grad_table || !grad_w0b || !grad_w1)
        return AC_E_INVALID_ARG;
    if (!(eps > 0.f) || M > 0xFFFFFFFFu / 7u) return AC_E_INVALID_ARG;
    if (M == 0) return AC_OK;
    return launch_sdf_backward(m, P, grad_centre, 7u * M, bound, scales, grad_table, grad_w0b, grad_w1, M, eps, grad_fd, (cudaStream_t)stream);
}

extern "C" uint64_t ac_nsr_sdf_backward_workspace_bytes(uint32_t n_points) { return (uint64_t)n_points * 32u * sizeof(float); }

extern "C" int ac_nsr_sdf_backward_fused_ws(const ac_nsr_model* m, const float* x, const float* grad_out, uint32_t B, float bound, const float* scales,
                                            float* grad_table, float* grad_w0b, float* grad_w1, void* workspace, uint64_t workspace_bytes,
                                            void* stream) {
    if (!m || !m->embeddings || !m->offsets || !m->mlp_blob || !x || !grad_out || !scales || !grad_table || !grad_w0b || !grad_w1)
        return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    return launch_sdf_backward(m, x, grad_out, B, bound, scales, grad_table, grad_w0b, grad_w1, 0u, 0.f, nullptr, (cudaStream_t)stream, workspace,
                               workspace_bytes);
}

extern "C" int ac_nsr_sdf_backward_stencil_ws(const ac_nsr_model* m, const float* P, uint32_t M, float bound, float eps, const float* grad_centre,
                                              const float* grad_fd, const float* scales, float* grad_table, float* grad_w0b, float* grad_w1,
                                              void* workspace, uint64_t workspace_bytes, const void* feature_cache, void* stream) {
    if (!m || !m->embeddings || !m->offsets || !m->mlp_blob || !P || !grad_centre || !grad_fd || !scales || !grad_table || !grad_w0b || !grad_w1)
        return AC_E_INVALID_ARG;
    if (!(eps > 0.f) || M > 0xFFFFFFFFu / 7u) return AC_E_INVALID_ARG;
    if (M == 0) return AC_OK;
    if (feature_cache && ((uintptr_t)feature_cache & 15)) return AC_E_INVALID_ARG;
    return launch_sdf_backward(m, P, grad_centre, 7u * M, bound, scales, grad_table, grad_w0b, grad_w1, M, eps, grad_fd, (cudaStream_t)stream, workspace,
                               workspace_bytes, feature_cache);
}
