// nsr_render_tc.cu -- the fused Instant-NSR render core, tensor-core edition (sm_100a).
//
// Persistent kernel: one CTA of 28 warps (7 groups x 4 warps, <= 73 registers) per SM, 512 TMEM columns.
// A warp owns a ray from box intersection to composited pixel (nsr_device.cuh); four warps form a
// GROUP whose 128 lanes are the 128 rows of one MMA tile.  Every dense layer that matters runs on the
// 5th-generation tensor cores (tcgen05.mma, kind::f16, fp32 accumulate in TMEM):
//
//   SDF layer 0   32 hash features -> 64 hidden      fp16x3: A = (hi, lo), B = (hi, lo); hi*hi + lo*hi + hi*lo,
//                                                    6 MMAs M128 N64 K16; relative error ~2^-21 (fp32-class)
//   colour 0      21 (padded 32) -> 64               fp16x3, 6 MMAs
//   colour 1      64 -> 64                           A = fp16(h1) single, B = (hi, lo): 8 MMAs (colour tolerates 2^-11)
//
// A rows are written by the lane that owns the point (hash-encode -> split -> st.shared.v4 into the
// UMMA K-major no-swizzle layout: 16-byte chunks, 2 KB chunk stride), one elected thread issues the MMAs
// and commits to the group's mbarrier, every lane reads its accumulator row back with tcgen05.ld.
// The raw-xyz columns of SDF layer 0 + bias are added in exact fp32 in the epilogue (they dominate the
// SDF and feed the +-0.005 finite differences), softplus(beta=100) costs two MUFU ops, and the 64 -> {1,16}
// second SDF layer and the 64 -> 3 colour head are short register dot products.
//
// fp16x3 instead of 3xTF32 halves the A-tile footprint (16 KB per group), which is what lets 7-8 groups
// share one SM: the kernel is bound by gather latency (ncu: long_scoreboard), so resident warps are a lever --
// up to the point where the register budget forces spills (8 groups = 64 registers is slower than 7 = 73).
// Groups are independent (own A tile, 64 TMEM columns, named barrier, mbarrier).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"
#include "tc05.cuh"
#include "nsr_tc_group.cuh"

using namespace acb;

// Tuning switches (scripts/bench_variants.sh builds and times the combinations on the GPU box).
#ifndef AC_SPECIALIZED_LEVELS
#define AC_SPECIALIZED_LEVELS 1     // 1: separate dense / hashed level bodies (fewer instructions, more code)
#endif
#ifndef AC_FULL_TAIL_LOOP
#define AC_FULL_TAIL_LOOP 0         // 1: 16-output SDF tail as a rolled loop (4x less code; ~2 % slower at 6-7 groups)
#endif

#ifndef AC_ENCODE_UNROLL
#define AC_ENCODE_UNROLL 1          // unroll factor of the 4-chunk encode loop
#endif
#ifndef AC_GROUPS
#define AC_GROUPS 7                 // 4-warp groups per CTA.  Measured ms/frame (B200, 256x256, 64+64): 4: 16.0  5: 16.0  6: 15.0  7: 14.5  8: 18.5 (64 regs, spills)
#endif
#define AC_PRAGMA(x) _Pragma(#x)
#define AC_UNROLL(n) AC_PRAGMA(unroll n)

namespace {

constexpr int kGroups = AC_GROUPS;                      // groups of 4 warps per CTA
constexpr int kWarpsTC = 4 * kGroups;           // 32 warps, 1024 threads
constexpr int kMaxT = 128;
constexpr uint32_t kTmemCols = 512;             // one 128x64 fp32 accumulator per group (8 x 64 = all of TMEM)

// fp32 weights of the register epilogues live in the constant bank: every lane reads the same element, the
// index is a compile-time constant after unrolling, so they fold into FFMA operands (c[bank][offset]) and cost
// neither a load instruction nor MIO-queue bandwidth.  Refreshed from the blob's OFF_EPI block by a 6 KB D2D
// copy on the launch stream before every launch.
// Re-entrancy: the bank holds kSlots copies and every kernel is instantiated once per slot (the slot must be a
// compile-time constant for the operand folding).  A launch leases the next slot round-robin, makes its stream
// wait for the event of the slot's previous user, copies, launches and records the slot's event (SlotLease at the
// end of this file): launches of different models on different streams never share a copy.
#ifndef AC_CW_SLOTS
#define AC_CW_SLOTS 4
#endif
constexpr int kSlots = AC_CW_SLOTS;
__constant__ float c_w[kSlots][EPI_FLOATS];
constexpr int W_XB = EPI_XB, W_W1T = EPI_W1T, W_B1 = EPI_B1, W_C2T = EPI_C2T;
#define CW(i) c_w[SLOT][(i)]

// dynamic shared memory map (bytes)
constexpr size_t SM_LEVELS = 0;
constexpr size_t SM_B = (SM_LEVELS + kLevels * sizeof(LevelMeta) + 127) / 128 * 128;
constexpr uint32_t B_W0_HI = 0, B_W0_LO = 4096, B_C0_HI = 8192, B_C0_LO = 12288, B_C1_HI = 16384, B_C1_LO = 24576, B_BYTES = 32768;
constexpr size_t SM_A = SM_B + B_BYTES;                                   // per group 16 KB: hi [4 chunks] | lo [4 chunks]
constexpr size_t SM_ROWS = SM_A + (size_t)kGroups * 16384;                // per warp: depths, sdf, scratch (3 x 128 floats)
constexpr size_t SM_BARS = SM_ROWS + (size_t)kWarpsTC * 3 * kMaxT * 4;
constexpr size_t SM_TOTAL = SM_BARS + kGroups * 8 + 16;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

struct RenderParamsTC {
    const float2* table;
    const int32_t* offsets;
    const float* blob;
    const float* variance;
    float S;
    uint32_t H;
    ac_nsr_render_args a;
    float* eik_partial;
};

// Epilogue of the SDF network for this thread's accumulator row: + raw-xyz columns + bias (exact fp32),
// softplus, second layer 64 -> {1, 16}.
template <int SLOT, bool FULL>
__device__ __forceinline__ void sdf_tail(uint32_t tmem_row, float x, float y, float z, float (&out)[FULL ? 16 : 1]) {
#pragma unroll
    for (int o = 0; o < (FULL ? 16 : 1); ++o) out[o] = CW(W_B1 + o);
#pragma unroll
    for (int qtr = 0; qtr < 4; ++qtr) {
        float acc[16];
        tc05::tmem_ld16(tmem_row + qtr * 16, acc);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int j = qtr * 16 + jj;
            const float lin = fmaf(CW(W_XB + 4 * j), x, fmaf(CW(W_XB + 4 * j + 1), y, fmaf(CW(W_XB + 4 * j + 2), z, CW(W_XB + 4 * j + 3))));
            const float h = softplus100_mufu(acc[jj] + lin);
            if (FULL) {
#pragma unroll
                for (int o = 0; o < 16; ++o) out[o] = fmaf(CW(W_W1T + j * 16 + o), h, out[o]);
            } else {
                out[0] = fmaf(CW(W_W1T + j * 16), h, out[0]);
            }
        }
    }
}
template <int SLOT>
__device__ __noinline__ float sdf_tail_scalar(uint32_t tmem_row, float x, float y, float z) {
    float o[1];
    sdf_tail<SLOT, false>(tmem_row, x, y, z, o);
    return o[0];
}

// Encode -> A tile -> tcgen05.mma -> epilogue.  Must be called by all 128 threads of the group, the same
// number of times.
template <int SLOT, bool FULL>
__device__ __forceinline__ void group_sdf_eval(Group& g, const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                               float bound, float x, float y, float z, float (&out)[FULL ? 16 : 1]) {
    encode_to_tile(g.a + g.row * 16, table, lv, bound, x, y, z, g.std_layout);
    group_mma_round(g, [&] { issue_k32_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + B_W0_HI, g.b_s + B_W0_LO); });
    if constexpr (FULL) {
#if AC_FULL_TAIL_LOOP
#pragma unroll
        for (int o = 0; o < 16; ++o) out[o] = CW(W_B1 + o);
#pragma unroll 1
        for (int qtr = 0; qtr < 4; ++qtr) {
            float acc[16];
            tc05::tmem_ld16(g.tmem + qtr * 16, acc);
            const float* __restrict__ xb = c_w[SLOT] + W_XB + 64 * qtr;
            const float* __restrict__ w1 = c_w[SLOT] + W_W1T + 256 * qtr;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float lin = fmaf(xb[4 * jj], x, fmaf(xb[4 * jj + 1], y, fmaf(xb[4 * jj + 2], z, xb[4 * jj + 3])));
                const float h = softplus100_mufu(acc[jj] + lin);
#pragma unroll
                for (int o = 0; o < 16; ++o) out[o] = fmaf(w1[jj * 16 + o], h, out[o]);
            }
        }
#else
        sdf_tail<SLOT, true>(g.tmem, x, y, z, out);
#endif
    } else {
        out[0] = sdf_tail_scalar<SLOT>(g.tmem, x, y, z);
    }
}

// Colour MLP 21 -> 64 -> 64 -> 3 (models/instant_nsr.py:644-663) for the group's 128 samples, after layer 0's MMA has
// been committed: relu -> layer 1 (fp16 activations x (hi, lo) weights) -> relu -> 64 -> 3 head -> sigmoid.
// VIEW: use_viewdirs=True (:564-569,646-650) -- the 16 SH coefficients of the ray direction enter layer 0 as a per-RAY bias
// bias[64] = C0[:, sh columns] sh(d) (the direction is the same for every sample of a ray), added before the relu.
template <int SLOT, bool VIEW = false>
__device__ __forceinline__ void group_color_rest(Group& g, float (&rgb)[3], const float* __restrict__ bias = nullptr) {
    // relu -> fp16 -> A tile of layer 1 (K = 64 = 8 chunks, fills the whole 16 KB region)
#pragma unroll
    for (int qtr = 0; qtr < 4; ++qtr) {
        float acc[16];
        tc05::tmem_ld16(g.tmem + qtr * 16, acc);
        if constexpr (VIEW) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) acc[jj] += __ldg(bias + qtr * 16 + jj);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint4 pk;
            pk.x = tc05::pack_f16x2(fmaxf(acc[8 * half + 0], 0.f), fmaxf(acc[8 * half + 1], 0.f));
            pk.y = tc05::pack_f16x2(fmaxf(acc[8 * half + 2], 0.f), fmaxf(acc[8 * half + 3], 0.f));
            pk.z = tc05::pack_f16x2(fmaxf(acc[8 * half + 4], 0.f), fmaxf(acc[8 * half + 5], 0.f));
            pk.w = tc05::pack_f16x2(fmaxf(acc[8 * half + 6], 0.f), fmaxf(acc[8 * half + 7], 0.f));
            *reinterpret_cast<uint4*>(g.a + (2 * qtr + half) * 2048 + g.row * 16) = pk;
        }
    }
    group_mma_round(g, [&] {
        constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
#pragma unroll
        for (uint32_t s = 0; s < 4; ++s) {           // K = 64 = 4 x (K=16)
            const uint64_t ad = tc05::smem_desc(g.a_s + s * 4096u, 2048u, 128u);
            const uint64_t bh = tc05::smem_desc(g.b_s + B_C1_HI + s * 2048u, 1024u, 128u);
            const uint64_t bl = tc05::smem_desc(g.b_s + B_C1_LO + s * 2048u, 1024u, 128u);
            tc05::mma_f16(g.tmem & 0xFFFFu, ad, bh, idesc, s);
            tc05::mma_f16(g.tmem & 0xFFFFu, ad, bl, idesc, 1u);
        }
    });
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
    for (int qtr = 0; qtr < 4; ++qtr) {
        float acc[16];
        tc05::tmem_ld16(g.tmem + qtr * 16, acc);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const float h2 = fmaxf(acc[jj], 0.f);
            const int j = qtr * 16 + jj;
            o0 = fmaf(CW(W_C2T + 4 * j), h2, o0); o1 = fmaf(CW(W_C2T + 4 * j + 1), h2, o1); o2 = fmaf(CW(W_C2T + 4 * j + 2), h2, o2);
        }
    }
    rgb[0] = sigmoidf(o0); rgb[1] = sigmoidf(o1); rgb[2] = sigmoidf(o2);
}

// cin = (x, y, z, nx, ny, nz, 15 geometry features); all 128 threads of the group call it together.
template <int SLOT, bool VIEW = false>
__device__ __forceinline__ void group_color_eval(Group& g, const float (&cin)[24], float (&rgb)[3], const float* __restrict__ bias = nullptr) {
    // layer 0: K = 32 (21 inputs + zero pad), fp16x3
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        if (c < 3) {
            tc05::split_f16x2(cin[8 * c + 0], cin[8 * c + 1], hi.x, lo.x);
            tc05::split_f16x2(cin[8 * c + 2], cin[8 * c + 3], hi.y, lo.y);
            tc05::split_f16x2(cin[8 * c + 4], cin[8 * c + 5], hi.z, lo.z);
            tc05::split_f16x2(cin[8 * c + 6], cin[8 * c + 7], hi.w, lo.w);
        } else {
            hi = make_uint4(0u, 0u, 0u, 0u); lo = hi;
        }
        *reinterpret_cast<uint4*>(g.a + c * 2048 + g.row * 16) = hi;
        *reinterpret_cast<uint4*>(g.a + 8192 + c * 2048 + g.row * 16) = lo;
    }
    group_mma_round(g, [&] { issue_k32_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + B_C0_HI, g.b_s + B_C0_LO); });
    group_color_rest<SLOT, VIEW>(g, rgb, bias);
}

template <int SLOT, bool VIEW = false, bool SKIP = false>      // SKIP: ac_nsr_render_args.skip_masked (its own instantiation: the
__global__ void __launch_bounds__(kWarpsTC * 32, 1) nsr_render_tc_kernel(const RenderParamsTC p) {   // default kernel's code is untouched)
    extern __shared__ __align__(1024) unsigned char smem[];
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + SM_LEVELS);
    unsigned char* bt = smem + SM_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroups);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp >> 2;

    // ---- one-time staging: fp32 epilogue weights, level table, fp16 weight tiles, barriers, TMEM ----
    {
        const float* blob = p.blob;
        if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(p.offsets, threadIdx.x, p.S, p.H, 3);
        for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
            const int n = i >> 5, k = i & 31;
            stage_b_tile(bt + B_W0_HI, bt + B_W0_LO, n, k, __ldg(blob + OFF_W0 + n * kSdfInPad + 3 + k));
            stage_b_tile(bt + B_C0_HI, bt + B_C0_LO, n, k, k < kColInPad ? __ldg(blob + OFF_C0 + n * kColInPad + k) : 0.f);
        }
        for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
            const int n = i >> 6, k = i & 63;
            stage_b_tile(bt + B_C1_HI, bt + B_C1_LO, n, k, __ldg(blob + OFF_C1 + n * kHidden + k));
        }
        if (threadIdx.x == 0) {
            for (int gI = 0; gI < kGroups; ++gI) tc05::mbar_init(bars + gI, 1);
            tc05::fence_mbar_init();
        }
        if (warp == 0) tc05::tmem_alloc<kTmemCols>(tmem_slot);
        tc05::fence_proxy_async_smem();
        tc05::fence_before_sync();
        __syncthreads();
        tc05::fence_after_sync();
    }
    const uint32_t tmem_base = *tmem_slot;

    Group g;
    g.a = smem + SM_A + (size_t)group * 16384;
    g.a_s = tc05::smem_u32(g.a);
    g.b_s = tc05::smem_u32(bt);
    g.bar = bars + group;
    g.phase = 0;
    g.row = (warp & 3) * 32 + lane;
    g.tmem = tmem_base + (uint32_t)group * 64u + ((uint32_t)((warp & 3) * 32) << 16);
    g.bar_id = 1 + group;
    {
        bool ok = true;
        for (int l = 0; l < kLevels; ++l) ok = ok && (lv[l].hashed == (l < 5 ? 0u : 1u));
        g.std_layout = ok;
    }

    float* zs = reinterpret_cast<float*>(smem + SM_ROWS) + warp * 3 * kMaxT;   // sorted depths
    float* sdfs = zs + kMaxT;                                                  // their SDF
    float* ta = sdfs + kMaxT;                                                  // scratch (alpha, then cdf)
    float* fd = sdfs;                                                          // 192 finite-difference values (render core)

    const float bound = p.a.bound;
    const float2* __restrict__ table = p.table;
    const int N0 = (int)p.a.num_steps;
    const int rounds = (int)p.a.upsample_steps / 16;
    const int Ttot = N0 + 16 * rounds;
    const float inv_s = clampf(expf(p.variance[0] * 10.0f), 1e-6f, 1e6f);
    const float eps = 0.005f * (1.0f - p.a.normal_epsilon_ratio);
    const float car = p.a.cos_anneal_ratio;
    const uint32_t n_quads = (p.a.n_rays + 3) / 4;
    const bool staged = p.a.z_in != nullptr;       // sampling done by the host pipeline (warp path)

    // A full launch gives a CTA kGroups adjacent quads per iteration; a small one deals quads to CTAs first, so that a
    // 512-ray shard of a training patch runs on 128 SMs with one busy group each instead of on 19 full CTAs.
    const bool dense_map = n_quads >= gridDim.x * kGroups;
    for (uint32_t quad = dense_map ? blockIdx.x * kGroups + group : blockIdx.x + gridDim.x * group; quad < n_quads;
         quad += gridDim.x * kGroups) {
        const uint32_t ray_raw = quad * 4 + (warp & 3);
        const bool ray_ok = ray_raw < p.a.n_rays;
        const uint32_t ray = ray_ok ? ray_raw : p.a.n_rays - 1;       // padding warps recompute the last ray
        Ray r;
        r.ox = p.a.rays_o[3 * ray + 0]; r.oy = p.a.rays_o[3 * ray + 1]; r.oz = p.a.rays_o[3 * ray + 2];
        r.dx = p.a.rays_d[3 * ray + 0]; r.dy = p.a.rays_d[3 * ray + 1]; r.dz = p.a.rays_d[3 * ray + 2];
        float near, far;
        if (p.a.near_far_in) { near = p.a.near_far_in[2 * ray]; far = p.a.near_far_in[2 * ray + 1]; }
        else ray_box(r, bound, near, far);
        const float span = far - near;
        const float sample_dist = span / (float)N0;
        if (staged) {
            for (int k = lane; k < Ttot; k += 32) zs[k] = p.a.z_in[(size_t)ray * Ttot + k];
            __syncwarp();
        }

        // ---- coarse samples (:155-174) and their SDF (:178) ----
        for (int k0 = 0; k0 < N0 && !staged; k0 += 32) {
            const int k = min(k0 + lane, N0 - 1);
            float z = near + span * linspace01(k, N0);
            if (p.a.jitter) z = z + (p.a.jitter[(size_t)ray * N0 + k] - 0.5f) * sample_dist;
            if (rounds > 0) {
                float x, y, zz;
                ray_point(r, z, x, y, zz);
                float o[1];
                group_sdf_eval<SLOT, false>(g, table, lv, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
                                      clampf(zz, -bound, bound), o);
                sdfs[k] = o[0];
            }
            zs[k] = z;          // duplicate lanes write identical values
        }
        __syncwarp();

        // ---- importance rounds (:182-184), merged in place ----
        int T = N0;
        for (int i = 0; i < rounds && !staged; ++i) {
            float z_new; int below, above;
            importance_round(r, zs, sdfs, ta, ta, T, (float)(64 << i), lane, z_new, below, above);
            float s_new = 0.0f;
            if (i + 1 < rounds) {                       // uniform across the launch: all 128 threads take it
                const float zq = __shfl_sync(0xffffffffu, z_new, lane & 15);     // lanes 16..31 mirror 0..15
                float x, y, zz;
                ray_point(r, zq, x, y, zz);
                float o[1];
                group_sdf_eval<SLOT, false>(g, table, lv, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
                                      clampf(zz, -bound, bound), o);
                s_new = o[0];
            }
            int pos_old[4], pos_new;
            merge_positions(zs, T, z_new, lane, pos_old, pos_new);
            float zo[4], so[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = lane + 32 * q;
                zo[q] = k < T ? zs[k] : 0.f;
                so[q] = k < T ? sdfs[k] : 0.f;
            }
            __syncwarp();                               // every lane holds its elements: safe to scatter in place
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (lane + 32 * q < T) { zs[pos_old[q]] = zo[q]; sdfs[pos_old[q]] = so[q]; }
            if (lane < 16) { zs[pos_new] = z_new; sdfs[pos_new] = s_new; }
            __syncwarp();
            T += 16;
        }

        if (p.a.rgb == nullptr) {        // sampling-only launch (training path): the sorted depths are the result
            if (ray_ok)
                for (int k = lane; k < Ttot; k += 32) p.a.z_vals[(size_t)ray * Ttot + k] = zs[k];
            __syncwarp();
            continue;
        }

        // ---- render core (:186-299), 32 section samples at a time in depth order ----
        // Per block: the 6 x 32 finite-difference points (:687-704) first, lanes packed as (sample, direction):
        // the six +-eps neighbours of a sample sit within 0.003 of each other, so on every level coarser than
        // ~10 they share a grid cell and the warp's gather collapses to a few cache lines.  Their SDF values
        // are parked in the scratch rows (sdfs/ta are free now); then the block is shaded (lane = sample) and
        // the transmittance is carried across blocks with a shuffle scan.
        float carry = 1.0f;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_nx = 0.f, acc_ny = 0.f, acc_nz = 0.f;
        float acc_w = 0.f, acc_d = 0.f, eik_num = 0.f, eik_den = 0.f;
        for (int k0 = 0; k0 < Ttot; k0 += 32) {
            const int nS = min(32, Ttot - k0);
            // skip_masked: when the warp's mask is zero on all 32 samples of this block for all four rays of the group (one
            // barrier with an OR reduction), nothing is evaluated -- every alpha of the block is zero anyway
            bool dead = false;
            if constexpr (SKIP) {
                const bool on = k0 + lane < Ttot && p.a.alpha_mask[(size_t)ray * Ttot + k0 + lane] != 0.0f;
                dead = !tc05::named_bar_or(g.bar_id, 128, on);
            }
            for (int q0 = 0; !dead && q0 < 6 * nS; q0 += 32) {
                const int q = min(q0 + lane, 6 * nS - 1);
                const int sI = q / 6, dir = q - 6 * sI;
                const int k = k0 + sI;
                float px, py, pz;
                if (p.a.pts_in) {
                    const float* q3 = p.a.pts_in + 3 * ((size_t)ray * Ttot + k);
                    px = q3[0]; py = q3[1]; pz = q3[2];
                } else {
                    const float zk = zs[k];
                    const float zmid = k < Ttot - 1 ? zk + 0.5f * (zs[k + 1] - zk) : zk;
                    ray_point(r, zmid, px, py, pz);
                }
                px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
                const float e = (dir & 1) ? -eps : eps;
                const int ax = dir >> 1;
                const float qx = ax == 0 ? clampf(px + e, -bound, bound) : px;
                const float qy = ax == 1 ? clampf(py + e, -bound, bound) : py;
                const float qz = ax == 2 ? clampf(pz + e, -bound, bound) : pz;
                float o[1];
                group_sdf_eval<SLOT, false>(g, table, lv, bound, qx, qy, qz, o);
                if (q0 + lane < 6 * nS) fd[q] = o[0];
            }
            __syncwarp();

            const bool live = k0 + lane < Ttot;
            const int k = min(k0 + lane, Ttot - 1);
            const float zk = zs[k];
            const float delta = k < Ttot - 1 ? zs[k + 1] - zk : sample_dist;
            float px, py, pz;
            if (p.a.pts_in) {
                const float* q3 = p.a.pts_in + 3 * ((size_t)ray * Ttot + k);
                px = q3[0]; py = q3[1]; pz = q3[2];
            } else {
                ray_point(r, k < Ttot - 1 ? zk + 0.5f * delta : zk, px, py, pz);
            }
            px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
            float cin[24], sdf0 = 0.f, gn = 1.0f;
            if (!dead) {
                float o16[16];
                group_sdf_eval<SLOT, true>(g, table, lv, bound, px, py, pz, o16);
                sdf0 = o16[0];
#pragma unroll
                for (int q = 0; q < 15; ++q) cin[6 + q] = o16[1 + q];
            }
            if (dead) {
                cin[3] = cin[4] = cin[5] = 0.f;
            } else {
                const float* f6 = fd + 6 * (k - k0);
                const float gx = 0.5f * (f6[0] - f6[1]) / eps, gy = 0.5f * (f6[2] - f6[3]) / eps, gz = 0.5f * (f6[4] - f6[5]) / eps;
                gn = sqrtf(gx * gx + gy * gy + gz * gz);
                const float inv = 1e-5f + gn;
                cin[3] = gx / inv; cin[4] = gy / inv; cin[5] = gz / inv;
            }
            cin[0] = px; cin[1] = py; cin[2] = pz;
            cin[21] = cin[22] = cin[23] = 0.f;
            __syncwarp();                                // fd[] consumed before the next block overwrites it
            float col[3] = {0.f, 0.f, 0.f};
            if (!p.a.opacity_only && !dead)              // launch-uniform: the trainer's frozen-net pass only reads weight_sum
                group_color_eval<SLOT, VIEW>(g, cin, col, VIEW ? p.a.c0_ray_bias + 64 * (size_t)ray : nullptr);
            const float nx = cin[3], ny = cin[4], nz = cin[5];
            const float cosv = r.dx * nx + r.dy * ny + r.dz * nz;
            const float it = -(softplus100(-cosv * 0.5f + 0.5f) * (1.0f - car) + softplus100(-cosv) * car);
            const float hs = it * delta * 0.5f;
            const float c0 = sigmoidf((sdf0 - hs) * inv_s), c1 = sigmoidf((sdf0 + hs) * inv_s);
            float alpha = clampf((c0 - c1 + 1e-5f) / (c0 + 1e-5f), 0.0f, 1.0f);
            if (p.a.alpha_mask) alpha = alpha * p.a.alpha_mask[(size_t)ray * Ttot + k];
            if (!live) alpha = 0.0f;
            const float pn = sqrtf(px * px + py * py + pz * pz);
            if (live && !dead && pn < 1.2f) { eik_num += (gn - 1.0f) * (gn - 1.0f); eik_den += 1.0f; }

            float blk;
            const float tr = warp_excl_prod(live ? (1.0f - alpha + 1e-7f) : 1.0f, lane, blk) * carry;
            carry *= blk;
            const float w = alpha * tr;
            if (live) {
                acc_r += w * col[0]; acc_g += w * col[1]; acc_b += w * col[2];
                acc_nx += w * nx; acc_ny += w * ny; acc_nz += w * nz;
                acc_w += w;
                acc_d += w * clampf((zk - near) / span, 0.0f, 1.0f);
                if (ray_ok) {
                    const size_t s = (size_t)ray * Ttot + k;
                    if (p.a.weights) p.a.weights[s] = w;
                    if (p.a.pts_alpha) p.a.pts_alpha[s] = alpha;
                    if (p.a.z_vals) p.a.z_vals[s] = zk;
                    if (p.a.pts_color) { p.a.pts_color[3 * s] = col[0]; p.a.pts_color[3 * s + 1] = col[1]; p.a.pts_color[3 * s + 2] = col[2]; }
                }
            }
        }
        acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
        acc_nx = warp_sum(acc_nx); acc_ny = warp_sum(acc_ny); acc_nz = warp_sum(acc_nz);
        acc_w = warp_sum(acc_w); acc_d = warp_sum(acc_d);
        eik_num = warp_sum(eik_num); eik_den = warp_sum(eik_den);
        if (lane == 0 && ray_ok) {
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
        __syncwarp();
    }

    // ---- teardown ----
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<kTmemCols>(tmem_base);
}

// NeRFNetwork.forward_sdf on a flat point list with the same tensor-core group machinery as the render kernel
// (thread = point, four warps = one 128-row MMA tile): 3.7 M points per training patch, 2x the SIMT kernel's rate.
// Staging is the render kernel's prologue restricted to what the SDF network needs.
template <int SLOT>
__global__ void __launch_bounds__(kWarpsTC * 32, 1) forward_sdf_tc_kernel(const float2* __restrict__ table, const int32_t* __restrict__ offsets,
                                                                          const float* __restrict__ blob, float S, uint32_t H,
                                                                          const float* __restrict__ x, float* __restrict__ out, uint32_t B, float bound,
                                                                          const uint32_t stencil_M, const float eps, float* __restrict__ out_fd,
                                                                          unsigned char* __restrict__ feat_cache, const uint32_t packed) {
    // feat_cache (optional): every 128-point tile's encoded A operand (32 features as fp16 hi | lo chunks, 16 KB, the layout the
    // MMA reads) is also written to feat_cache + tile * 16 KB, so the backward of the same points loads it back with
    // coalesced 16-byte accesses instead of repeating the 128 gathers per point.
    // stencil_M > 0: x holds M section points and the B = 7 M evaluated points are generated here -- block 0 the points
    // themselves (16 outputs -> out [M,16]), blocks 1..6 their +-eps neighbours along x, y, z re-clamped to the bound
    // (signed distance only -> out_fd [6,M]): the finite-difference stencil of NeRFNetwork.gradient (:683-704).
    extern __shared__ __align__(1024) unsigned char smem[];
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + SM_LEVELS);
    unsigned char* bt = smem + SM_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroups);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, group = warp >> 2;
    if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(offsets, threadIdx.x, S, H, 3);
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x)
        stage_b_tile(bt + B_W0_HI, bt + B_W0_LO, i >> 5, i & 31, __ldg(blob + OFF_W0 + (i >> 5) * kSdfInPad + 3 + (i & 31)));
    if (threadIdx.x == 0) {
        for (int gI = 0; gI < kGroups; ++gI) tc05::mbar_init(bars + gI, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<kTmemCols>(tmem_slot);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    Group g;
    g.a = smem + SM_A + (size_t)group * 16384;
    g.a_s = tc05::smem_u32(g.a);
    g.b_s = tc05::smem_u32(bt);
    g.bar = bars + group;
    g.phase = 0;
    g.row = (warp & 3) * 32 + lane;
    g.tmem = tmem_base + (uint32_t)group * 64u + ((uint32_t)((warp & 3) * 32) << 16);
    g.bar_id = 1 + group;
    {
        bool ok = true;
        for (int l = 0; l < kLevels; ++l) ok = ok && (lv[l].hashed == (l < 5 ? 0u : 1u));
        g.std_layout = ok;
    }
    const uint32_t per_cta = kGroups * 128;
    if (packed) {
        // Stencil mode, M a multiple of 128: a group takes 128 section points and evaluates their 7 x 128 stencil points in seven
        // rounds.  Rounds 0..5 are the neighbours with a warp's lanes packed as (sample, direction) -- the six +-eps neighbours of a
        // sample sit within 0.01 of each other, so on every level coarser than ~10 they share a grid cell and the warp's gather
        // collapses to a few cache lines (the render kernel's arrangement, :687-704) -- round 6 the points themselves.  Outputs and
        // feature-cache rows go where the point-major numbering b = block * M + sample puts them, so the backward is unchanged.
        const int wq = warp & 3;
        // 128-sample blocks are dealt to CTAs first, then to a CTA's groups: a small launch still spreads over every SM
        for (uint32_t u = blockIdx.x + gridDim.x * (uint32_t)group; u < (stencil_M >> 7); u += gridDim.x * kGroups) {    // group-uniform
            const uint32_t s0 = (u << 7) + 32u * (uint32_t)wq;
#pragma unroll 1
            for (int k = 0; k < 7; ++k) {
                uint32_t blk = 0, smp = s0 + (uint32_t)lane;
                if (k < 6) {
                    const uint32_t q = 32u * (uint32_t)k + (uint32_t)lane;
                    smp = s0 + q / 6u; blk = q % 6u + 1u;
                }
                float px = x[3 * (size_t)smp], py = x[3 * (size_t)smp + 1], pz = x[3 * (size_t)smp + 2];
                if (blk) {
                    const float e = (blk & 1) ? eps : -eps;  // blocks 1,3,5 = +eps, 2,4,6 = -eps
                    const uint32_t ax = (blk - 1) >> 1;
                    if (ax == 0) px = clampf(px + e, -bound, bound);
                    else if (ax == 1) py = clampf(py + e, -bound, bound);
                    else pz = clampf(pz + e, -bound, bound);
                }
                if (k == 6) {
                    float o16[16];
                    group_sdf_eval<SLOT, true>(g, table, lv, bound, px, py, pz, o16);
                    float4* dst = reinterpret_cast<float4*>(out + 16 * (size_t)smp);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
                } else {
                    float o1[1];
                    group_sdf_eval<SLOT, false>(g, table, lv, bound, px, py, pz, o1);
                    out_fd[(size_t)(blk - 1) * stencil_M + smp] = o1[0];
                }
                if (feat_cache) {
                    const size_t b = (size_t)blk * stencil_M + smp;
                    unsigned char* dst = feat_cache + (b >> 7) * 16384 + (b & 127) * 16;
                    const unsigned char* src = g.a + g.row * 16;
#pragma unroll
                    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + c * 2048) = *reinterpret_cast<const uint4*>(src + c * 2048);
                }
            }
        }
    } else
    for (uint32_t base = blockIdx.x * per_cta; base < B; base += gridDim.x * per_cta) {      // uniform trip count per CTA
        const uint32_t b = base + threadIdx.x;
        const bool valid = b < B;
        const uint32_t bb = valid ? b : B - 1;           // padding threads re-evaluate the last point (every thread joins the MMA round)
        float px, py, pz;
        uint32_t blk = 0, smp = bb;
        if (stencil_M) {
            blk = bb / stencil_M; smp = bb - blk * stencil_M;
            px = x[3 * (size_t)smp]; py = x[3 * (size_t)smp + 1]; pz = x[3 * (size_t)smp + 2];
            if (blk) {
                const float e = (blk & 1) ? eps : -eps;  // blocks 1,3,5 = +eps, 2,4,6 = -eps
                const uint32_t ax = (blk - 1) >> 1;
                if (ax == 0) px = clampf(px + e, -bound, bound);
                else if (ax == 1) py = clampf(py + e, -bound, bound);
                else pz = clampf(pz + e, -bound, bound);
            }
        } else {
            px = x[3 * (size_t)bb]; py = x[3 * (size_t)bb + 1]; pz = x[3 * (size_t)bb + 2];
        }
        // the group's 128 threads take the same branch: all 16 outputs when the group's first point is a section point
        const uint32_t g_first = base + (uint32_t)(group * 128);
        if (!stencil_M || g_first < stencil_M) {
            float o16[16];
            group_sdf_eval<SLOT, true>(g, table, lv, bound, px, py, pz, o16);
            if (valid) {
                if (blk == 0) {
                    float4* dst = reinterpret_cast<float4*>(out + 16 * (size_t)smp);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(o16[4 * q], o16[4 * q + 1], o16[4 * q + 2], o16[4 * q + 3]);
                } else {
                    out_fd[(size_t)(blk - 1) * stencil_M + smp] = o16[0];
                }
            }
        } else {
            float o1[1];
            group_sdf_eval<SLOT, false>(g, table, lv, bound, px, py, pz, o1);
            if (valid) out_fd[(size_t)(blk - 1) * stencil_M + smp] = o1[0];
        }
        if (feat_cache && g_first < B) {                  // this thread's row of the tile: 8 chunks of 16 bytes, 2 KB apart
            unsigned char* dst = feat_cache + (size_t)(g_first >> 7) * 16384 + g.row * 16;
            const unsigned char* src = g.a + g.row * 16;
#pragma unroll
            for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(dst + c * 2048) = *reinterpret_cast<const uint4*>(src + c * 2048);
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<kTmemCols>(tmem_base);
}

// Unit-test kernel for the tensor-core layer alone: feats [128,32] (fp32) x W0[:,3:35]^T -> acc [128,64] (fp16x3).
__global__ void __launch_bounds__(128, 1) debug_tc_layer_kernel(const float* __restrict__ feats, const float* __restrict__ blob,
                                                                float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a = smem;                      // 16 KB
    unsigned char* bhi = smem + 16384;            // 4 KB
    unsigned char* blo = bhi + 4096;              // 4 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(blo + 4096);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, row = threadIdx.x;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) stage_b_tile(bhi, blo, i >> 5, i & 31, blob[OFF_W0 + (i >> 5) * kSdfInPad + 3 + (i & 31)]);
    for (int c = 0; c < 4; ++c) {
        uint4 hi, lo;
        const float* f = feats + row * 32 + 8 * c;
        tc05::split_f16x2(f[0], f[1], hi.x, lo.x); tc05::split_f16x2(f[2], f[3], hi.y, lo.y);
        tc05::split_f16x2(f[4], f[5], hi.z, lo.z); tc05::split_f16x2(f[6], f[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(a + c * 2048 + row * 16) = hi;
        *reinterpret_cast<uint4*>(a + 8192 + c * 2048 + row * 16) = lo;
    }
    if (threadIdx.x == 0) { tc05::mbar_init(bar, 1); tc05::fence_mbar_init(); }
    if (warp == 0) tc05::tmem_alloc<64>(slot);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        issue_k32_x3(tmem, tc05::smem_u32(a), tc05::smem_u32(bhi), tc05::smem_u32(blo));
        tc05::mma_commit(bar);
    }
    tc05::mbar_wait(bar, 0);
    tc05::fence_after_sync();
    for (int qtr = 0; qtr < 4; ++qtr) {
        float acc[16];
        tc05::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + qtr * 16, acc);
        for (int j = 0; j < 16; ++j) out[row * 64 + qtr * 16 + j] = acc[j];
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64>(tmem);
}

}  // namespace

#include "nsr_render_st.cuh"

namespace {
// ---- constant-bank slot leases -------------------------------------------------------------------------------
struct SlotTable {
    cudaEvent_t ev[kSlots] = {};
    bool has_ev[kSlots] = {};
    unsigned next = 0;
};
SlotTable g_slots[64];
std::mutex g_slot_mu;

// Holds the mutex from lease to event record: the three host calls (copy, launch, record) of one launch are not
// interleaved with another thread's, so "the slot's event" always covers the slot's latest user.
struct SlotLease {
    std::lock_guard<std::mutex> lock;
    cudaStream_t st;
    SlotTable* tab = nullptr;
    int slot = 0;
    bool capturing = false;
    int rc = AC_OK;
    SlotLease(const float* blob, cudaStream_t stream) : lock(g_slot_mu), st(stream) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { rc = acb::cuda_fail(); return; }
        tab = &g_slots[dev];
        slot = (int)(tab->next++ % kSlots);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        capturing = cs != cudaStreamCaptureStatusNone;       // inside a capture the graph's own edges order the slot's users
        if (!capturing && tab->has_ev[slot] && cudaStreamWaitEvent(st, tab->ev[slot], 0) != cudaSuccess) { rc = acb::cuda_fail(); return; }
        if (cudaMemcpyToSymbolAsync(c_w, blob + OFF_EPI, EPI_FLOATS * sizeof(float), (size_t)slot * EPI_FLOATS * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = acb::cuda_fail();
    }
    void launched() {
        if (capturing || !tab) return;
        if (!tab->has_ev[slot]) {
            if (cudaEventCreateWithFlags(&tab->ev[slot], cudaEventDisableTiming) != cudaSuccess) return;
            tab->has_ev[slot] = true;
        }
        cudaEventRecord(tab->ev[slot], st);
    }
};

#if AC_CW_SLOTS == 4
#define AC_SLOT_SWITCH(slot, CALL)        \
    switch (slot) {                       \
        case 0: { CALL(0); } break;       \
        case 1: { CALL(1); } break;       \
        case 2: { CALL(2); } break;       \
        default: { CALL(3); } break;      \
    }
#elif AC_CW_SLOTS == 1                    /* development builds: one instantiation, 4x faster to compile */
#define AC_SLOT_SWITCH(slot, CALL) { CALL(0); }
#else
#error "AC_CW_SLOTS must be 1 or 4"
#endif
}  // namespace

namespace acb {
int launch_render_tc(const ac_nsr_model* m, const ac_nsr_render_args* a, cudaStream_t st) {
    RenderParamsTC p;
    p.table = reinterpret_cast<const float2*>(m->embeddings);
    p.offsets = m->offsets; p.blob = m->mlp_blob; p.variance = m->variance;
    p.S = m->log2_per_level_scale; p.H = m->base_resolution;
    p.a = *a;
    p.eik_partial = reinterpret_cast<float*>(a->workspace);
    // AC_RENDER_IMPL=tc5: the round-1 kernel (one point per lane, no stencil sharing), kept for A/B tests.
    const char* impl = getenv("AC_RENDER_IMPL");          // read per launch so that one process can compare both kernels
    const bool use_v5 = a->c0_ray_bias || !(impl && impl[0] == 's' && impl[1] == 't');      // WIP: the stencil kernel is opt-in (AC_RENDER_IMPL=st) until it is the faster one
    const uint32_t n_quads = (a->n_rays + 3) / 4;
    const uint32_t sms = (uint32_t)acb::sm_count();
    if (!use_v5) {
        // The stencil kernel keeps its epilogue weights in shared memory: no constant-bank slot, nothing shared between
        // launches.  One 4-ray quad per group; quads are dealt to CTAs first, so a small launch (a 512-ray shard of a
        // training patch) spreads over all SMs with one busy group each instead of filling 1/4 of the SMs.
        const uint32_t grid = n_quads < sms ? n_quads : sms;
        ACB_SET_MAX_SMEM(nsr_render_st_kernel, SS_TOTAL);
        nsr_render_st_kernel<<<grid, kWarpsS * 32, SS_TOTAL, st>>>(p);
        return acb::launched();
    }
#if defined(AC_TEX_MIN_SCALE)
    {   // experiment: linear float2 texture over the table (standard 16-level layout: 6 119 857 entries)
        static cudaTextureObject_t tex = 0; static const void* tex_ptr = nullptr;
        if (tex_ptr != m->embeddings) {
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = const_cast<float*>(m->embeddings);
            rd.res.linear.desc = cudaCreateChannelDesc<float2>(); rd.res.linear.sizeInBytes = (size_t)6119857 * 8;
            cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) return acb::cuda_fail();
            tex_ptr = m->embeddings;
        }
        if (cudaMemcpyToSymbolAsync(c_table_tex, &tex, sizeof(tex), 0, cudaMemcpyHostToDevice, st) != cudaSuccess) return acb::cuda_fail();
    }
#endif
    SlotLease lease(m->mlp_blob, st);
    if (lease.rc) return lease.rc;
    const uint32_t grid = n_quads < sms ? n_quads : sms;
    if (a->c0_ray_bias) {                 // use_viewdirs: separate instantiation, the default kernel's registers are untouched
#define AC_CALL(S)                                                                \
    ACB_SET_MAX_SMEM((nsr_render_tc_kernel<S, true>), SM_TOTAL);                  \
    nsr_render_tc_kernel<S, true><<<grid, kWarpsTC * 32, SM_TOTAL, st>>>(p)
        AC_SLOT_SWITCH(lease.slot, AC_CALL)
#undef AC_CALL
    } else if (a->skip_masked && a->alpha_mask && a->rgb) {      // warped inference that keeps only the image
#define AC_CALL(S)                                                                \
    ACB_SET_MAX_SMEM((nsr_render_tc_kernel<S, false, true>), SM_TOTAL);           \
    nsr_render_tc_kernel<S, false, true><<<grid, kWarpsTC * 32, SM_TOTAL, st>>>(p)
        AC_SLOT_SWITCH(lease.slot, AC_CALL)
#undef AC_CALL
    } else {
#define AC_CALL(S)                                                                \
    ACB_SET_MAX_SMEM((nsr_render_tc_kernel<S, false>), SM_TOTAL);                 \
    nsr_render_tc_kernel<S, false><<<grid, kWarpsTC * 32, SM_TOTAL, st>>>(p)
        AC_SLOT_SWITCH(lease.slot, AC_CALL)
#undef AC_CALL
    }
    const int rc = acb::launched();
    lease.launched();
    return rc;
}

int launch_forward_sdf_tc(const ac_nsr_model* m, const float* x, float* out, uint32_t B, float bound, cudaStream_t st, uint32_t stencil_M,
                          float eps, float* out_fd, void* feat_cache) {
    const uint32_t per_cta = kGroups * 128;
    // packed stencil rounds (see the kernel) when the section points fill whole 128-row tiles and the machine; AC_STENCIL_FWD=flat: A/B
    static const bool flat = [] { const char* e = getenv("AC_STENCIL_FWD"); return e && e[0] == 'f'; }();
    const uint32_t packed = (stencil_M && !flat && (stencil_M & 127u) == 0u) ? 1u : 0u;
    const uint32_t want = packed ? (stencil_M >> 7) : (B + per_cta - 1) / per_cta;
    const uint32_t grid = want < (uint32_t)acb::sm_count() ? want : (uint32_t)acb::sm_count();
    SlotLease lease(m->mlp_blob, st);
    if (lease.rc) return lease.rc;
#define AC_CALL(S)                                                                                                              \
    ACB_SET_MAX_SMEM(forward_sdf_tc_kernel<S>, SM_TOTAL);                                                                       \
    forward_sdf_tc_kernel<S><<<grid, kWarpsTC * 32, SM_TOTAL, st>>>(reinterpret_cast<const float2*>(m->embeddings), m->offsets, \
        m->mlp_blob, m->log2_per_level_scale, m->base_resolution, x, out, B, bound, stencil_M, eps, out_fd,           \
        reinterpret_cast<unsigned char*>(feat_cache), packed)
    AC_SLOT_SWITCH(lease.slot, AC_CALL)
#undef AC_CALL
    const int rc = acb::launched();
    lease.launched();
    return rc;
}
}  // namespace acb

extern "C" int ac_nsr_debug_tc_layer(const float* feats, const float* blob, float* out, void* stream) {
    if (!feats || !blob || !out) return AC_E_INVALID_ARG;
    const size_t smem = 16384 + 8192 + 32;
    debug_tc_layer_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(feats, blob, out);
    return acb::launched();
}
