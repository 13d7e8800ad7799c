// nsr_tc_group.cuh -- the 4-warp "group" machinery shared by the tensor-core kernels of the Instant-NSR path
// (nsr_render_tc.cu, nsr_shade_tc.cu): four warps = the 128 rows of one tcgen05.mma tile, lane = row; every thread writes
// its own A row (K-major no-swizzle layout: 16-byte chunks of 8 fp16, 2 KB chunk stride, hi tile then lo tile), one
// elected thread issues the MMAs and commits to the group's mbarrier, every thread reads its accumulator row back with
// tcgen05.ld.  Each translation unit gets its own copy (anonymous namespace).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsr_device.cuh"
#include "tc05.cuh"

using namespace acb;

namespace {

struct Group {
    unsigned char* a;     // this group's A region (generic pointer)
    uint32_t a_s;         // ... and its shared-space address
    uint32_t b_s;         // shared-space address of the weight tiles (layout B_*: W0 | C0 | C1, each hi then lo)
    uint64_t* bar;
    uint32_t phase;
    uint32_t tmem;        // accumulator address of this thread's row (lane field set), column 0 of the group
    uint32_t bar_id;
    int row;              // 0..127 inside the group
    bool std_layout;      // offsets table has the reference layout (5 dense + 11 power-of-two hashed levels)
};

// Make this thread's A-row stores visible to the tensor core, rendezvous the group, let one thread issue,
// wait for the commit.  `issue` runs in exactly one thread.
template <class Issue>
__device__ __forceinline__ void group_mma_round(Group& g, Issue issue) {
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    tc05::named_bar_sync(g.bar_id, 128);
    if (g.row == 0) {
        tc05::fence_after_sync();
        issue();
        tc05::mma_commit(g.bar);
    }
    tc05::mbar_wait(g.bar, g.phase);
    g.phase ^= 1u;
    tc05::fence_after_sync();
}

// D[128x64] = A(hi,lo)[128x32] * B(hi,lo)[64x32]^T with the three fp16 partial products.
__device__ __forceinline__ void issue_k32_x3(uint32_t tmem_d, uint32_t a_s, uint32_t bhi_s, uint32_t blo_s) {
    constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
#pragma unroll
    for (uint32_t s = 0; s < 2; ++s) {               // K = 32 = 2 x (K=16): chunks 2s, 2s+1
        const uint64_t ah = tc05::smem_desc(a_s + s * 4096u, 2048u, 128u);
        const uint64_t al = tc05::smem_desc(a_s + 8192u + s * 4096u, 2048u, 128u);
        const uint64_t bh = tc05::smem_desc(bhi_s + s * 2048u, 1024u, 128u);
        const uint64_t bl = tc05::smem_desc(blo_s + s * 2048u, 1024u, 128u);
        tc05::mma_f16(tmem_d, ah, bh, idesc, s);
        tc05::mma_f16(tmem_d, al, bh, idesc, 1u);
        tc05::mma_f16(tmem_d, ah, bl, idesc, 1u);
    }
}

// Weight W[n][k] (n < 64) -> fp16 (hi, lo) tiles in the UMMA K-major layout: chunk k/8, row n, element k%8.
__device__ __forceinline__ void stage_b_tile(unsigned char* bhi, unsigned char* blo, int n, int k, float w) {
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const int at = (k >> 3) * 1024 + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(bhi + at) = h;
    *reinterpret_cast<__half*>(blo + at) = l;
}

// Hash-encode one point (features only) into this thread's A-tile row (hi chunks at arow + c * 2048, lo chunks LO bytes
// further).  ONE copy of this code serves every
// call site of a kernel (__noinline__): with 28 warps in seven independent phases the instruction cache, not the
// issue slots, was the first thing to saturate when it was inlined four times.
template <uint32_t LO = 8192u>
__device__ __forceinline__ void store_chunk(unsigned char* arow, int c, float2 f0, float2 f1, float2 f2, float2 f3) {
    uint4 hi, lo;
    tc05::split_f16x2(f0.x, f0.y, hi.x, lo.x);
    tc05::split_f16x2(f1.x, f1.y, hi.y, lo.y);
    tc05::split_f16x2(f2.x, f2.y, hi.z, lo.z);
    tc05::split_f16x2(f3.x, f3.y, hi.w, lo.w);
    *reinterpret_cast<uint4*>(arow + c * 2048) = hi;
    *reinterpret_cast<uint4*>(arow + LO + c * 2048) = lo;
}

// Any offsets table (cold path): level kind decided per level at run time.
template <uint32_t LO = 8192u>
__device__ __noinline__ void encode_to_tile_generic(unsigned char* arow, const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                                    float u, float v, float w) {
#pragma unroll 1
    for (int c = 0; c < 4; ++c)
        store_chunk<LO>(arow, c, grid_level_3d(table, lv[4 * c + 0], u, v, w), grid_level_3d(table, lv[4 * c + 1], u, v, w),
                    grid_level_3d(table, lv[4 * c + 2], u, v, w), grid_level_3d(table, lv[4 * c + 3], u, v, w));
}

// `std_layout`: levels 0..4 dense, 5..15 hashed with power-of-two sizes -- the reference's only configuration
// (models/instant_nsr.py:503-512); verified once per CTA from the offsets table.
template <uint32_t LO = 8192u>
__device__ __noinline__ void encode_to_tile(unsigned char* arow, const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                            float bound, float x, float y, float z, bool std_layout) {
    const float two_b = 2.0f * bound;
    const float u = (x + bound) / two_b, v = (y + bound) / two_b, w = (z + bound) / two_b;
    if ((u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f)) {       // hashencoder.cu:94-119: zeros
        const float2 z2 = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) store_chunk<LO>(arow, c, z2, z2, z2, z2);
        return;
    }
    if (!std_layout) { encode_to_tile_generic<LO>(arow, table, lv, u, v, w); return; }
    store_chunk<LO>(arow, 0, grid_level_3d_k<false>(table, lv[0], u, v, w), grid_level_3d_k<false>(table, lv[1], u, v, w),
                grid_level_3d_k<false>(table, lv[2], u, v, w), grid_level_3d_k<false>(table, lv[3], u, v, w));
    store_chunk<LO>(arow, 1, grid_level_3d_k<false>(table, lv[4], u, v, w), grid_level_3d_k<true>(table, lv[5], u, v, w),
                grid_level_3d_k<true>(table, lv[6], u, v, w), grid_level_3d_k<true>(table, lv[7], u, v, w));
#pragma unroll 1
    for (int c = 2; c < 4; ++c)
        store_chunk<LO>(arow, c, grid_level_3d_k<true>(table, lv[4 * c + 0], u, v, w), grid_level_3d_k<true>(table, lv[4 * c + 1], u, v, w),
                    grid_level_3d_k<true>(table, lv[4 * c + 2], u, v, w), grid_level_3d_k<true>(table, lv[4 * c + 3], u, v, w));
}

}  // namespace
