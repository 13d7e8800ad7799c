// nsr_render_tc.cu -- the fused Instant-NSR render core, tensor-core edition (sm_100a).
//
// Persistent kernel, one CTA of 16 warps per SM.  A warp still owns a ray from box intersection
// to composited pixel (nsr_device.cuh), but the SDF network's first layer -- 32 hash features ->
// 64 hidden units, 84 % of the network's multiply-adds -- runs on the 5th-generation tensor cores:
//
//   * four warps form a GROUP; one lane = one sample point = one row of a 128 x 32 A tile.  Each
//     lane hash-encodes its point, splits the 32 fp32 features into tf32 (hi, lo) pairs and stores
//     them into the group's A tiles in the UMMA K-major no-swizzle layout (16-byte row chunks,
//     conflict-free st.shared.v4).
//   * one elected thread issues 12 tcgen05.mma.kind::tf32 (M128 N64 K8; 4 K-steps x {hi*hi, lo*hi,
//     hi*lo} = "3xTF32", error ~2^-22) against the weight tiles resident in shared memory; the fp32
//     accumulator lives in the group's 64 TMEM columns; tcgen05.commit -> mbarrier.
//   * every lane pulls its accumulator row back with tcgen05.ld, adds the raw-xyz columns and the
//     bias in exact fp32 (the xyz part dominates the SDF and feeds the +-0.005 finite differences),
//     applies softplus(beta=100) and finishes the 64 -> {1,16} second layer in registers.
//
// Groups are independent (own A tiles, TMEM columns, named barrier, mbarrier), so while one group
// waits on its MMA the other three keep the LSU / FMA / MUFU pipes busy with gathers and epilogues.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"
#include "tc05.cuh"

using namespace acb;

namespace {

constexpr int kGroups = 4;                      // groups of 4 warps per CTA
constexpr int kWarpsTC = 4 * kGroups;           // 16 warps, 512 threads
constexpr int kMaxT = 128;
constexpr int kFeat = 32;                       // hash features = K of the tensor-core layer
constexpr uint32_t kTmemCols = 64 * kGroups;    // 256 columns: one 128x64 fp32 accumulator per group

// dynamic shared memory map (bytes)
constexpr size_t SM_BLOB = 0;                                             // packed fp32 weights (SIMT part)
constexpr size_t SM_LEVELS = SM_BLOB + BLOB_FLOATS * 4;                   // 16 x LevelMeta
constexpr size_t SM_B = (SM_LEVELS + kLevels * sizeof(LevelMeta) + 127) / 128 * 128;   // B_hi 8 KB, B_lo 8 KB
constexpr size_t SM_A = SM_B + 2 * 64 * kFeat * 4;                        // per group: A_hi 16 KB, A_lo 16 KB
constexpr size_t SM_ROWS = SM_A + (size_t)kGroups * 2 * 128 * kFeat * 4;  // per warp: 4 rows x 128 floats
constexpr size_t SM_BARS = SM_ROWS + (size_t)kWarpsTC * 4 * kMaxT * 4;    // mbarriers + tmem slot
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

struct Group {
    float* a_hi;          // this group's A tiles, [8 chunks][128 rows][4 floats]
    float* a_lo;
    uint32_t a_hi_s, a_lo_s, b_hi_s, b_lo_s;   // shared-space addresses for the descriptors
    uint64_t* bar;
    uint32_t phase;
    uint32_t tmem;        // TMEM address of this thread's accumulator row (lane field set), column 0 of the group
    uint32_t tmem_d;      // accumulator base for the MMA (lane 0)
    uint32_t bar_id;
    int row;              // 0..127 inside the group
};

// Hash-encode (features only) -> A tiles -> tcgen05.mma -> epilogue.  Must be called by all 128
// threads of the group, the same number of times.
template <bool FULL>
__device__ __forceinline__ void group_sdf_eval(Group& g, const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                               const float* __restrict__ sw, float bound, float x, float y, float z,
                                               float (&out)[FULL ? 16 : 1]) {
    {
        const float two_b = 2.0f * bound;
        const float u = (x + bound) / two_b, v = (y + bound) / two_b, w = (z + bound) / two_b;
        const bool oob = (u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {               // 2 levels = 4 features = one 16-byte chunk of the row
            float2 f0 = make_float2(0.f, 0.f), f1 = f0;
            if (!oob) {
                f0 = grid_level_3d(table, lv[2 * c], u, v, w);
                f1 = grid_level_3d(table, lv[2 * c + 1], u, v, w);
            }
            float4 hi, lo;
            tc05::split_tf32(f0.x, hi.x, lo.x); tc05::split_tf32(f0.y, hi.y, lo.y);
            tc05::split_tf32(f1.x, hi.z, lo.z); tc05::split_tf32(f1.y, hi.w, lo.w);
            *reinterpret_cast<float4*>(g.a_hi + c * 512 + g.row * 4) = hi;
            *reinterpret_cast<float4*>(g.a_lo + c * 512 + g.row * 4) = lo;
        }
    }
    tc05::fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
    tc05::fence_before_sync();           // orders this thread's earlier tcgen05.ld before the next MMA
    tc05::named_bar_sync(g.bar_id, 128);
    if (g.row == 0) {
        tc05::fence_after_sync();
        constexpr uint32_t idesc = tc05::idesc_tf32(128, 64);
#pragma unroll
        for (uint32_t s = 0; s < 4; ++s) {           // K = 32 = 4 x (K=8): chunks 2s, 2s+1
            const uint64_t ah = tc05::smem_desc(g.a_hi_s + s * 4096u, 2048u, 128u);
            const uint64_t al = tc05::smem_desc(g.a_lo_s + s * 4096u, 2048u, 128u);
            const uint64_t bh = tc05::smem_desc(g.b_hi_s + s * 2048u, 1024u, 128u);
            const uint64_t bl = tc05::smem_desc(g.b_lo_s + s * 2048u, 1024u, 128u);
            tc05::mma_tf32(g.tmem_d, ah, bh, idesc, s > 0 ? 1u : 0u);
            tc05::mma_tf32(g.tmem_d, al, bh, idesc, 1u);
            tc05::mma_tf32(g.tmem_d, ah, bl, idesc, 1u);
        }
        tc05::mma_commit(g.bar);
    }
    tc05::mbar_wait(g.bar, g.phase);
    g.phase ^= 1u;
    tc05::fence_after_sync();

#pragma unroll
    for (int o = 0; o < (FULL ? 16 : 1); ++o) out[o] = sw[OFF_B1 + o];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float acc[32];
        tc05::tmem_ld32(g.tmem + half * 32, acc);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = half * 32 + jj;
            const float4 wx = *reinterpret_cast<const float4*>(sw + OFF_W0 + j * kSdfInPad);   // (wx, wy, wz, .)
            const float lin = fmaf(wx.x, x, fmaf(wx.y, y, fmaf(wx.z, z, sw[OFF_B0 + j])));
            const float h = softplus100_mufu(acc[jj] + lin);
            if (FULL) {
                const float4* __restrict__ w1 = reinterpret_cast<const float4*>(sw + OFF_W1T + j * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w4 = w1[q];
                    out[4 * q + 0] = fmaf(w4.x, h, out[4 * q + 0]);
                    out[4 * q + 1] = fmaf(w4.y, h, out[4 * q + 1]);
                    out[4 * q + 2] = fmaf(w4.z, h, out[4 * q + 2]);
                    out[4 * q + 3] = fmaf(w4.w, h, out[4 * q + 3]);
                }
            } else {
                out[0] = fmaf(sw[OFF_W1T + j * 16], h, out[0]);
            }
        }
    }
}

__global__ void __launch_bounds__(kWarpsTC * 32, 1) nsr_render_tc_kernel(const RenderParamsTC p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sw = reinterpret_cast<float*>(smem + SM_BLOB);
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + SM_LEVELS);
    float* b_hi = reinterpret_cast<float*>(smem + SM_B);
    float* b_lo = b_hi + 64 * kFeat;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroups);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp >> 2;

    // ---- one-time staging: fp32 blob, level table, tf32 weight tiles, barriers, TMEM ----
    {
        const float4* src = reinterpret_cast<const float4*>(p.blob);
        float4* dst = reinterpret_cast<float4*>(sw);
        for (int i = threadIdx.x; i < BLOB_FLOATS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(p.offsets, threadIdx.x, p.S, p.H, 3);
        // B tiles: W0[n][3 + k] for n < 64, k < 32 -> [chunk k/4][n][k%4], hi and lo parts
        for (int i = threadIdx.x; i < 64 * kFeat; i += blockDim.x) {
            const int n = i / kFeat, k = i % kFeat;
            float hi, lo;
            tc05::split_tf32(__ldg(p.blob + OFF_W0 + n * kSdfInPad + 3 + k), hi, lo);
            const int at = (k >> 2) * 256 + n * 4 + (k & 3);
            b_hi[at] = hi;
            b_lo[at] = lo;
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
    g.a_hi = reinterpret_cast<float*>(smem + SM_A) + (size_t)group * 2 * 128 * kFeat;
    g.a_lo = g.a_hi + 128 * kFeat;
    g.a_hi_s = tc05::smem_u32(g.a_hi); g.a_lo_s = tc05::smem_u32(g.a_lo);
    g.b_hi_s = tc05::smem_u32(b_hi); g.b_lo_s = tc05::smem_u32(b_lo);
    g.bar = bars + group;
    g.phase = 0;
    g.row = (warp & 3) * 32 + lane;
    g.tmem_d = tmem_base + (uint32_t)group * 64u;
    g.tmem = g.tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    g.bar_id = 1 + group;

    float* zs = reinterpret_cast<float*>(smem + SM_ROWS) + warp * 4 * kMaxT;
    float* sdfs = zs + kMaxT;
    float* ta = sdfs + kMaxT;
    float* tb = ta + kMaxT;

    const float bound = p.a.bound;
    const float2* __restrict__ table = p.table;
    const int N0 = (int)p.a.num_steps;
    const int rounds = (int)p.a.upsample_steps / 16;
    const int Ttot = N0 + 16 * rounds;
    const float inv_s = clampf(expf(p.variance[0] * 10.0f), 1e-6f, 1e6f);
    const float eps = 0.005f * (1.0f - p.a.normal_epsilon_ratio);
    const float car = p.a.cos_anneal_ratio;
    const uint32_t n_quads = (p.a.n_rays + 3) / 4;

    for (uint32_t quad = blockIdx.x * kGroups + group; quad < n_quads; quad += gridDim.x * kGroups) {
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
        const bool staged = p.a.z_in != nullptr;       // sampling done by the host pipeline (warp path)
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
                group_sdf_eval<false>(g, table, lv, sw, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
                                      clampf(zz, -bound, bound), o);
                sdfs[k] = o[0];
            }
            zs[k] = z;          // duplicate lanes write identical values
        }
        __syncwarp();

        // ---- importance rounds (:182-184) ----
        int T = N0;
        for (int i = 0; i < rounds && !staged; ++i) {
            float z_new; int below, above;
            importance_round(r, zs, sdfs, ta, tb, T, (float)(64 << i), lane, z_new, below, above);
            float s_new = 0.0f;
            if (i + 1 < rounds) {                       // uniform across the launch: all 128 threads take it
                const float zq = __shfl_sync(0xffffffffu, z_new, lane & 15);     // lanes 16..31 mirror 0..15
                float x, y, zz;
                ray_point(r, zq, x, y, zz);
                float o[1];
                group_sdf_eval<false>(g, table, lv, sw, bound, clampf(x, -bound, bound), clampf(y, -bound, bound),
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

        // ---- render core (:186-299) ----
        // Section samples are processed in halves of 64.  For each half, first the 6 x 64 finite-
        // difference points (:687-704) are evaluated with lanes packed as (sample, direction): the
        // six +-eps neighbours of a sample sit within 0.003 of each other, so on every level coarser
        // than ~10 they fall in the same grid cell and the warp's gather collapses to a handful of
        // cache lines (the L1 tag stage, one line per cycle, is what bounds this kernel).  Their SDF
        // values are parked in three scratch rows; then the half's two 32-sample blocks are shaded
        // in depth order (lane = sample), carrying the transmittance across blocks.
        if (zs != reinterpret_cast<float*>(smem + SM_ROWS) + warp * 4 * kMaxT) {   // park the depths in row 0
            float* row0 = reinterpret_cast<float*>(smem + SM_ROWS) + warp * 4 * kMaxT;
            for (int k = lane; k < Ttot; k += 32) row0[k] = zs[k];
            __syncwarp();
            zs = row0;
        }
        float* fd = zs + kMaxT;            // rows 1..3: 384 floats = 64 samples x 6 directions
        sdfs = zs + kMaxT; ta = zs + 2 * kMaxT; tb = zs + 3 * kMaxT;
        float carry = 1.0f;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_nx = 0.f, acc_ny = 0.f, acc_nz = 0.f;
        float acc_w = 0.f, acc_d = 0.f, eik_num = 0.f, eik_den = 0.f;
        for (int h0 = 0; h0 < Ttot; h0 += 64) {
            const int nS = min(64, Ttot - h0);
            for (int q0 = 0; q0 < 6 * nS; q0 += 32) {
                const int q = min(q0 + lane, 6 * nS - 1);
                const int sI = q / 6, dir = q - 6 * sI;
                const int k = h0 + sI;
                const float zk = zs[k];
                const float zmid = k < Ttot - 1 ? zk + 0.5f * (zs[k + 1] - zk) : zk;
                float px, py, pz;
                if (p.a.pts_in) {
                    const float* q3 = p.a.pts_in + 3 * ((size_t)ray * Ttot + k);
                    px = q3[0]; py = q3[1]; pz = q3[2];
                } else ray_point(r, zmid, px, py, pz);
                px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
                const float e = (dir & 1) ? -eps : eps;
                const int ax = dir >> 1;
                const float qx = ax == 0 ? clampf(px + e, -bound, bound) : px;
                const float qy = ax == 1 ? clampf(py + e, -bound, bound) : py;
                const float qz = ax == 2 ? clampf(pz + e, -bound, bound) : pz;
                float o[1];
                group_sdf_eval<false>(g, table, lv, sw, bound, qx, qy, qz, o);
                if (q0 + lane < 6 * nS) fd[q] = o[0];
            }
            __syncwarp();
            for (int k0 = h0; k0 < h0 + nS; k0 += 32) {
            const bool live = k0 + lane < Ttot;
            const int k = min(k0 + lane, Ttot - 1);
            const float zk = zs[k];
            const float delta = k < Ttot - 1 ? zs[k + 1] - zk : sample_dist;
            const float zmid = k < Ttot - 1 ? zk + 0.5f * delta : zk;
            float px, py, pz;
            if (p.a.pts_in) {
                const float* q3 = p.a.pts_in + 3 * ((size_t)ray * Ttot + k);
                px = q3[0]; py = q3[1]; pz = q3[2];
            } else ray_point(r, zmid, px, py, pz);
            px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
            float o16[16];
            group_sdf_eval<true>(g, table, lv, sw, bound, px, py, pz, o16);
            float gr[3];
            {
                const float* f6 = fd + 6 * (k - h0);
                gr[0] = 0.5f * (f6[0] - f6[1]) / eps;
                gr[1] = 0.5f * (f6[2] - f6[3]) / eps;
                gr[2] = 0.5f * (f6[4] - f6[5]) / eps;
            }
            const float gn = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
            const float inv = 1e-5f + gn;
            float nrm[3] = {gr[0] / inv, gr[1] / inv, gr[2] / inv};
            float cin[kColInPad], col[3];
            cin[0] = px; cin[1] = py; cin[2] = pz; cin[3] = nrm[0]; cin[4] = nrm[1]; cin[5] = nrm[2];
#pragma unroll
            for (int q = 0; q < 15; ++q) cin[6 + q] = o16[1 + q];
            cin[21] = cin[22] = cin[23] = 0.f;
            color_mlp(sw, cin, col);
            const float cosv = r.dx * nrm[0] + r.dy * nrm[1] + r.dz * nrm[2];
            const float it = -(softplus100(-cosv * 0.5f + 0.5f) * (1.0f - car) + softplus100(-cosv) * car);
            const float hs = it * delta * 0.5f;
            const float c0 = sigmoidf((o16[0] - hs) * inv_s), c1 = sigmoidf((o16[0] + hs) * inv_s);
            float alpha = clampf((c0 - c1 + 1e-5f) / (c0 + 1e-5f), 0.0f, 1.0f);
            if (p.a.alpha_mask) alpha = alpha * p.a.alpha_mask[(size_t)ray * Ttot + k];
            if (!live) alpha = 0.0f;
            const float pn = sqrtf(px * px + py * py + pz * pz);
            if (live && pn < 1.2f) { eik_num += (gn - 1.0f) * (gn - 1.0f); eik_den += 1.0f; }

            float blk;
            const float tr = warp_excl_prod(live ? (1.0f - alpha + 1e-7f) : 1.0f, lane, blk) * carry;
            carry *= blk;
            const float w = alpha * tr;
            if (live) {
                acc_r += w * col[0]; acc_g += w * col[1]; acc_b += w * col[2];
                acc_nx += w * nrm[0]; acc_ny += w * nrm[1]; acc_nz += w * nrm[2];
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
            }   // 32-sample block
            __syncwarp();
        }       // 64-sample half
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

// Unit-test kernel for the tensor-core layer alone: feats [128,32] (fp32) x W0[:,3:35]^T -> acc [128,64].
__global__ void __launch_bounds__(128, 1) debug_tc_layer_kernel(const float* __restrict__ feats, const float* __restrict__ blob,
                                                                float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);
    float* a_lo = a_hi + 128 * kFeat;
    float* b_hi = a_lo + 128 * kFeat;
    float* b_lo = b_hi + 64 * kFeat;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + 64 * kFeat);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
    for (int i = threadIdx.x; i < 64 * kFeat; i += blockDim.x) {
        const int n = i / kFeat, k = i % kFeat;
        float hi, lo;
        tc05::split_tf32(blob[OFF_W0 + n * kSdfInPad + 3 + k], hi, lo);
        b_hi[(k >> 2) * 256 + n * 4 + (k & 3)] = hi;
        b_lo[(k >> 2) * 256 + n * 4 + (k & 3)] = lo;
    }
    for (int c = 0; c < 8; ++c) {
        float4 hi, lo;
        const float* f = feats + row * kFeat + 4 * c;
        tc05::split_tf32(f[0], hi.x, lo.x); tc05::split_tf32(f[1], hi.y, lo.y);
        tc05::split_tf32(f[2], hi.z, lo.z); tc05::split_tf32(f[3], hi.w, lo.w);
        *reinterpret_cast<float4*>(a_hi + c * 512 + row * 4) = hi;
        *reinterpret_cast<float4*>(a_lo + c * 512 + row * 4) = lo;
    }
    if (threadIdx.x == 0) { tc05::mbar_init(bar, 1); tc05::fence_mbar_init(); }
    if (warp == 0) tc05::tmem_alloc<64>(slot);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc05::idesc_tf32(128, 64);
        for (uint32_t s = 0; s < 4; ++s) {
            const uint64_t ah = tc05::smem_desc(tc05::smem_u32(a_hi) + s * 4096u, 2048u, 128u);
            const uint64_t al = tc05::smem_desc(tc05::smem_u32(a_lo) + s * 4096u, 2048u, 128u);
            const uint64_t bh = tc05::smem_desc(tc05::smem_u32(b_hi) + s * 2048u, 1024u, 128u);
            const uint64_t bl = tc05::smem_desc(tc05::smem_u32(b_lo) + s * 2048u, 1024u, 128u);
            tc05::mma_tf32(tmem, ah, bh, idesc, s > 0 ? 1u : 0u);
            tc05::mma_tf32(tmem, al, bh, idesc, 1u);
            tc05::mma_tf32(tmem, ah, bl, idesc, 1u);
        }
        tc05::mma_commit(bar);
    }
    tc05::mbar_wait(bar, 0);
    tc05::fence_after_sync();
    for (int half = 0; half < 2; ++half) {
        float acc[32];
        tc05::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, acc);
        for (int j = 0; j < 32; ++j) out[row * 64 + half * 32 + j] = acc[j];
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64>(tmem);
    (void)lane;
}

}  // namespace

namespace acb {
int launch_render_tc(const ac_nsr_model* m, const ac_nsr_render_args* a, cudaStream_t st) {
    RenderParamsTC p;
    p.table = reinterpret_cast<const float2*>(m->embeddings);
    p.offsets = m->offsets; p.blob = m->mlp_blob; p.variance = m->variance;
    p.S = m->log2_per_level_scale; p.H = m->base_resolution;
    p.a = *a;
    p.eik_partial = reinterpret_cast<float*>(a->workspace);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(nsr_render_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL); attr = true; }
    const uint32_t n_quads = (a->n_rays + 3) / 4;
    const uint32_t want = (n_quads + kGroups - 1) / kGroups;
    const uint32_t grid = want < (uint32_t)acb::sm_count() ? want : (uint32_t)acb::sm_count();
    nsr_render_tc_kernel<<<grid, kWarpsTC * 32, SM_TOTAL, st>>>(p);
    return acb::launched();
}
}  // namespace acb

extern "C" int ac_nsr_debug_tc_layer(const float* feats, const float* blob, float* out, void* stream) {
    if (!feats || !blob || !out) return AC_E_INVALID_ARG;
    const size_t smem = (size_t)(2 * 128 * kFeat + 2 * 64 * kFeat) * 4 + 32;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(debug_tc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    debug_tc_layer_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(feats, blob, out);
    return acb::launched();
}
