// sd_ops.cu -- native forward path of the Stable-Diffusion UNet the SDS step evaluates without gradients
// (reference: models/diffusion.py:121-132 calls diffusers' UNet2DConditionModel).  sm_100a.
//
// Activations are fp32 NHWC ([B*H*W, C] row-major == the token matrix of the transformer blocks); every GEMM operand is
// fp16 and every contraction runs on the 5th-generation tensor cores:
//
//   ac_sd_gemm_f16      C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ group_bias[m / rows_per_group, N]) (+ residual[M,N])
//                       tcgen05.mma kind::f16, M128 x N128 x K16, fp32 accumulators in TMEM (128 columns).  Default main
//                       loop: operands by TMA (cp.async.bulk.tensor 4-D tiles, SWIZZLE_128B, zero-filled past M / N / K),
//                       full/empty mbarrier ring, one elected producer thread and one elected MMA thread; batched through
//                       blockIdx.z with a two-level (outer, inner) stride folded into the tensor map, so per-head slices of
//                       [B, L, heads*d] need no copies; split-K for weight-streaming shapes.  The first version (cp.async
//                       into the no-swizzle layout, one __syncthreads per k-tile) is kept behind AC_SD_GEMM=cpasync.
//   ac_sd_conv3x3_f16   the same kernel as an implicit GEMM: the A tile of k-tile (tap, channel block) is the activation
//                       window of the CTA's 128 output pixels shifted by the tap, one 4-D TMA box; no im2col buffer.
//   ac_sd_flash_attention_f16   softmax(scale Q K^T) V per (batch, head, 128 queries) with the scores kept on the SM.
//   producers           write the fp16 A operand of the next GEMM: GroupNorm(+SiLU) applied on the fly inside the
//                       normalise-and-cast / im2col producer (3x3 stride 2, nearest x2 up-sampling), LayerNorm -> fp16,
//                       GEGLU -> fp16, softmax -> fp16 (unfused attention path, head dims > 128), plain cast.
//
// HBM traffic is irrelevant at these sizes (the largest activation is 10 MB); the GEMMs are L2-bandwidth / tensor bound.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "tc05.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------------------------
#ifndef AC_GEMM_STAGES
#define AC_GEMM_STAGES 3          // 3 x 32 KB: two CTAs per SM.  2 stages (three CTAs per SM) measure the same (11.57 vs 11.48 ms of GEMM
#endif                            // per guidance call): the 128 x 128 tile's 64 flop per operand byte is capped by the L2 slices, not by occupancy
constexpr int BM = 128, BN = 128, BK = 64, STAGES = AC_GEMM_STAGES;
constexpr int TILE_BYTES = BM * BK * 2;                 // 16 KB per operand per stage
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 128;   // + barriers + TMEM slot
constexpr int GEMM_THREADS = 128;

struct GemmParams {
    const __half* A; const __half* W;
    const float* bias; const float* group_bias; const float* residual;
    void* C;
    int M, N, K;
    long long lda, ldw, ldc, ldr;
    int rows_per_group;
    int batch_inner;
    long long sAo, sAi, sWo, sWi, sCo, sCi;
    int out_f16;
    int splits;            // split-K slices (blockIdx.z, batch == 1 only): partial sums meet in fp32 atomics on a zeroed C
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 128 x 64 fp16 tile (rows row0.., columns k0..) of a row-major matrix -> K-major no-swizzle core-matrix layout:
// 16-byte chunk c of row r at c*2048 + r*16.  A warp instruction covers 8 rows x 4 chunks: every 32 B sector it
// touches is consumed whole, and the 8 lanes of a shared-memory phase hit 8 distinct 16 B bank groups.
__device__ __forceinline__ void load_tile(uint32_t dst, const __half* __restrict__ base, long long ld, int rows, int K, int row0, int k0,
                                          int warp, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int u = warp * 8 + i;
        const int r = (u >> 1) * 8 + (lane & 7);
        const int c = (u & 1) * 4 + (lane >> 3);
        const int row = row0 + r, k = k0 + c * 8;
        int bytes = (K - k) * 2;
        bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
        if (row >= rows) bytes = 0;
        const __half* src = bytes > 0 ? base + (long long)row * ld + k : base;
        cp_async_16(dst + c * 2048 + r * 16, src, bytes);
    }
}

// Epilogue shared by both main loops.  TMEM lane = output row.  The tile is parked in shared memory (the operand ring is
// idle now) in two halves of 64 columns (68-float pitch: the 16-byte stores of a quarter-warp land on distinct banks; 34 KB, so
// it fits the two-stage ring) so that the global side is row-contiguous: one warp instruction reads/writes 256 B of each of
// two output rows (bias, per-row-group bias, residual and the store / split-K reduction).
__device__ __forceinline__ void gemm_epilogue(const GemmParams& p, unsigned char* smem, uint32_t tmem, int m0, int n0, long long c_off, int split,
                                              int warp, int lane) {
    float* stage = reinterpret_cast<float*>(smem);
    constexpr int HALF = BN / 2, PITCH = HALF + 4;
    static_assert(BM * PITCH * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit the operand ring");
    const bool lead = split == 0;                        // split-K: the first slice carries bias / residual
    const int sub = lane >> 4, l16 = lane & 15;          // a warp instruction covers two rows x 64 columns
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        if (n0 + h * HALF >= p.N) break;                 // CTA-uniform: nothing of this half is inside the matrix
        {
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * HALF);
            float* srow = stage + (warp * 32 + lane) * PITCH;
#pragma unroll 2
            for (int q = 0; q < HALF / 16; ++q) {
                float acc[16];
                tc05::tmem_ld16(trow + q * 16, acc);
#pragma unroll
                for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(srow + q * 16)[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
            }
        }
        __syncthreads();
        const int n = n0 + h * HALF + l16 * 4;
        const int nv = p.N - n < 4 ? p.N - n : 4;            // valid columns of this lane's float4 (<= 0: none)
        const bool vec_ok = nv == 4 && (p.ldc & 3) == 0 && (c_off & 3) == 0;
        float bz[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias && lead)
            for (int j = 0; j < 4; ++j) if (j < nv) bz[j] = p.bias[n + j];
#pragma unroll 1
        for (int r = warp * 2 + sub; r < BM; r += GEMM_THREADS / 16) {
            const int row = m0 + r;
            if (row >= p.M || nv <= 0) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(stage + r * PITCH + l16 * 4);
            float v[4] = {a4.x + bz[0], a4.y + bz[1], a4.z + bz[2], a4.w + bz[3]};
            if (lead && p.group_bias) {
                const float* gb = p.group_bias + (long long)(row / p.rows_per_group) * p.N + n;
                for (int j = 0; j < 4; ++j) if (j < nv) v[j] += gb[j];
            }
            if (lead && p.residual) {
                const float* rs = p.residual + (long long)row * p.ldr + n;
                if (nv == 4 && (p.ldr & 3) == 0 && ((reinterpret_cast<uintptr_t>(rs) & 15) == 0)) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rs);
                    v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
                } else {
                    for (int j = 0; j < 4; ++j) if (j < nv) v[j] += rs[j];
                }
            }
            const long long at = c_off + (long long)row * p.ldc + n;
            if (p.splits > 1) {
                float* dst = reinterpret_cast<float*>(p.C) + at;
                for (int j = 0; j < 4; ++j) if (j < nv) atomicAdd(dst + j, v[j]);
            } else if (p.out_f16) {
                __half* dst = reinterpret_cast<__half*>(p.C) + at;
                if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
                    uint2 o; o.x = tc05::pack_f16x2(v[0], v[1]); o.y = tc05::pack_f16x2(v[2], v[3]);
                    *reinterpret_cast<uint2*>(dst) = o;
                } else {
                    for (int j = 0; j < 4; ++j) if (j < nv) dst[j] = __float2half_rn(v[j]);
                }
            } else {
                float* dst = reinterpret_cast<float*>(p.C) + at;
                if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                else for (int j = 0; j < 4; ++j) if (j < nv) dst[j] = v[j];
            }
        }
        __syncthreads();                                 // the staging rows are rewritten by the second half
    }
}

__global__ void __launch_bounds__(GEMM_THREADS) sd_gemm_kernel(const GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* free_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + STAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int split = p.splits > 1 ? (int)blockIdx.z : 0;
    const int zb = p.splits > 1 ? 0 : (int)blockIdx.z;
    const int zo = zb / p.batch_inner, zi = zb % p.batch_inner;
    const __half* A = p.A + zo * p.sAo + zi * p.sAi;
    const __half* W = p.W + zo * p.sWo + zi * p.sWi;
    const long long c_off = zo * p.sCo + zi * p.sCi;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) tc05::mbar_init(free_bar + s, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<BN>(tmem_slot);
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t smem_s = tc05::smem_u32(smem);

    const int KT_all = (p.K + BK - 1) / BK;
    const int kt_begin = (int)((long long)KT_all * split / p.splits), kt_end = (int)((long long)KT_all * (split + 1) / p.splits);
    const int KT = kt_end - kt_begin;                    // k-tiles of this CTA (>= 1: the host keeps splits <= KT_all)
    const int kbase = kt_begin * BK;
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            load_tile(smem_s + s * STAGE_BYTES, A, p.lda, p.M, p.K, m0, kbase + s * BK, warp, lane);
            load_tile(smem_s + s * STAGE_BYTES + TILE_BYTES, W, p.ldw, p.N, p.K, n0, kbase + s * BK, warp, lane);
        }
        cp_async_commit();
    }
    constexpr uint32_t idesc = tc05::idesc_f16(BM, BN);
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();                    // this thread's copies of tile kt have landed
        tc05::fence_proxy_async_smem();                 // ... and are visible to the tensor core's (async-proxy) reads
        tc05::fence_before_sync();
        __syncthreads();
        const int stage = kt % STAGES;
        if (tid == 0) {
            tc05::fence_after_sync();
            const uint32_t a_s = smem_s + stage * STAGE_BYTES, b_s = a_s + TILE_BYTES;
#pragma unroll
            for (uint32_t s = 0; s < BK / 16; ++s)
                tc05::mma_f16(tmem, tc05::smem_desc(a_s + s * 4096u, 2048u, 128u), tc05::smem_desc(b_s + s * 4096u, 2048u, 128u), idesc,
                              (kt | (int)s) != 0 ? 1u : 0u);
            tc05::mma_commit(free_bar + stage);          // arrives when the MMAs that read this stage are done
        }
        const int nk = kt + STAGES - 1;                  // refill the stage tile kt-1 used
        if (nk < KT) {
            if (kt >= 1) tc05::mbar_wait(free_bar + (kt - 1) % STAGES, (uint32_t)(((kt - 1) / STAGES) & 1));
            const int ns = nk % STAGES;
            load_tile(smem_s + ns * STAGE_BYTES, A, p.lda, p.M, p.K, m0, kbase + nk * BK, warp, lane);
            load_tile(smem_s + ns * STAGE_BYTES + TILE_BYTES, W, p.ldw, p.N, p.K, n0, kbase + nk * BK, warp, lane);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();
    tc05::mbar_wait(free_bar + (KT - 1) % STAGES, (uint32_t)(((KT - 1) / STAGES) & 1));     // the last commit covers every MMA
    tc05::fence_after_sync();

    gemm_epilogue(p, smem, tmem, m0, n0, c_off, split, warp, lane);
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<BN>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------
// TMA main loop (default).  The cp.async ring above keeps the LSU / L1 data pipe 70 % busy moving operands (ncu r01i);
// here one elected thread issues `cp.async.bulk.tensor` 4-D tile loads (128 rows x 64 halves, SWIZZLE_128B, out-of-bounds
// rows / columns zero-filled by the TMA unit) that complete on a per-stage "full" mbarrier, a second elected thread
// issues the tcgen05.mma's on SWIZZLE_128B K-major descriptors and releases stages through tcgen05.commit -> "empty"
// mbarrier: no __syncthreads and no LSU traffic in the loop.  Tensor dims: (K, rows, batch_inner, batch_outer).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc05::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc05::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// K-major SWIZZLE_128B operand tile (rows of 128 B, 8-row groups 1024 B apart, 1024 B aligned base).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

struct ConvGeom { int taps, cblocks, W, H; };             // taps = 0: plain GEMM

__global__ void __launch_bounds__(GEMM_THREADS) sd_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                                   const GemmParams p, const int a_zi, const int a_zo, const int w_zi, const int w_zo,
                                                                   const ConvGeom cv) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* done_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int split = p.splits > 1 ? (int)blockIdx.z : 0;
    const int zb = p.splits > 1 ? 0 : (int)blockIdx.z;
    const int zo = zb / p.batch_inner, zi = zb % p.batch_inner;
    const long long c_off = zo * p.sCo + zi * p.sCi;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { tc05::mbar_init(full_bar + s, 1); tc05::mbar_init(empty_bar + s, 1); }
        tc05::mbar_init(done_bar, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<BN>(tmem_slot);
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t smem_s = tc05::smem_u32(smem);

    const int KT_all = (p.K + BK - 1) / BK;
    const int kt_begin = (int)((long long)KT_all * split / p.splits), kt_end = (int)((long long)KT_all * (split + 1) / p.splits);
    const int KT = kt_end - kt_begin;

    if (warp == 0 && lane == 0) {                       // ---- TMA producer ----
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            if (kt >= STAGES) tc05::mbar_wait(empty_bar + s, (uint32_t)(((kt / STAGES) - 1) & 1));
            mbar_expect_tx(full_bar + s, STAGE_BYTES);
            const int k0 = (kt_begin + kt) * BK;
            if (cv.taps) {
                // implicit 3x3 convolution: k-tile -> (tap, 64-channel block); the A tile is the activation window of this
                // CTA's 128 output pixels shifted by the tap, fetched as one 4-D box (C, W, H, B) whose out-of-image part
                // the TMA unit zero-fills (= the conv's zero padding).  No im2col buffer exists.
                const int kt_g = kt_begin + kt, tap = kt_g / cv.cblocks, cb = kt_g - tap * cv.cblocks;
                const int pix = m0 / cv.W, b0 = pix / cv.H, h0 = pix - b0 * cv.H, x0 = m0 - pix * cv.W;   // x0 != 0: rows wider than a tile
                tma_load_4d(smem_s + s * STAGE_BYTES, &tmA, full_bar + s, cb * 64, x0 + tap % 3 - 1, h0 + tap / 3 - 1, b0);
            } else {
                tma_load_4d(smem_s + s * STAGE_BYTES, &tmA, full_bar + s, k0, m0, a_zi ? zi : 0, a_zo ? zo : 0);
            }
            tma_load_4d(smem_s + s * STAGE_BYTES + TILE_BYTES, &tmW, full_bar + s, k0, n0, w_zi ? zi : 0, w_zo ? zo : 0);
        }
    } else if (warp == 1 && lane == 0) {                // ---- MMA issuer ----
        constexpr uint32_t idesc = tc05::idesc_f16(BM, BN);
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            tc05::mbar_wait(full_bar + s, (uint32_t)((kt / STAGES) & 1));
            tc05::fence_after_sync();
            const uint64_t da = smem_desc_sw128(smem_s + s * STAGE_BYTES), db = smem_desc_sw128(smem_s + s * STAGE_BYTES + TILE_BYTES);
#pragma unroll
            for (uint32_t k4 = 0; k4 < BK / 16; ++k4)   // +32 B along K inside the 128 B swizzle atom = +2 in the address field
                tc05::mma_f16(tmem, da + 2ull * k4, db + 2ull * k4, idesc, (kt | (int)k4) != 0 ? 1u : 0u);
            tc05::mma_commit(empty_bar + s);
        }
        tc05::mma_commit(done_bar);
    }
    __syncwarp();
    tc05::mbar_wait(done_bar, 0u);
    tc05::fence_after_sync();
    __syncthreads();                                     // every warp has seen the last MMA complete before the ring is reused
    gemm_epilogue(p, smem, tmem, m0, n0, c_off, split, warp, lane);
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<BN>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}
// fp16 [bo][bi][rows][K] view with element strides (ld, s_i, s_o); a zero batch stride (operand shared by the batch)
// collapses that dimension (the kernel then always asks for coordinate 0).  Returns false if the driver refuses.
inline bool make_operand_map(CUtensorMap* map, const void* base, int K, int rows, long long ld, int bi, long long s_i, int bo, long long s_o,
                             int* use_zi, int* use_zo) {
    *use_zi = (bi > 1 && s_i != 0) ? 1 : 0;
    *use_zo = (bo > 1 && s_o != 0) ? 1 : 0;
    const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(*use_zi ? bi : 1), (cuuint64_t)(*use_zo ? bo : 1)};
    const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
    const cuuint64_t strides[3] = {row_bytes, *use_zi ? (cuuint64_t)s_i * 2 : row_bytes, *use_zo ? (cuuint64_t)s_o * 2 : row_bytes};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)BM, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------------------------
// Fused attention: softmax(scale * Q K^T) V for one (batch, head, 128-query tile) per CTA, without the [Lq, Lk] score
// matrix ever leaving the SM (the unfused path writes 1.07 GB of fp32 scores + 0.54 GB of fp16 probabilities per
// 64x64 self-attention layer).  Two passes over the key tiles:
//   pass 1   S = Q K_j^T on tcgen05 (accumulator in TMEM), thread = query row: running max m and sum l (online softmax
//            statistics only -- no output, so nothing has to be rescaled later);
//   pass 2   S again, P = exp(scale S - m) / l written as the fp16 SWIZZLE_128B A operand, O += P V_j on tcgen05 with the
//            accumulator resident in TMEM; O -> fp16 -> [B, Lq, heads*d].
// Recomputing S costs one extra (cheap, K = head_dim) MMA per tile and removes every read-modify-write of O.
// Operands arrive by TMA (Q once, K and V^T double-buffered one tile ahead); one elected thread issues TMA and MMA, all
// 128 threads do the softmax arithmetic; the tile loop is synchronous (two CTAs per SM overlap each other's latencies
// when head_dim <= 64).  V^T [B, heads*d, Lk] is what the value projection GEMM already produces (swapped operands).
// ------------------------------------------------------------------------------------------------------------------
struct FlashParams {
    __half* out;            // [B, Lq, ld_out], head h at column h*d
    long long ld_out;
    int Lq, Lk, d, dn;      // dn = d rounded up to 16 (N of the P V MMA)
    float scale;
};

__device__ __forceinline__ float ex2_approx(float x) {       // 2^x, one MUFU.EX2 (2^-inf = 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int KA>           // 64-wide K atoms covering the head dim (1: d <= 64, 2: d <= 128)
__global__ void __launch_bounds__(128) sd_flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                            const __grid_constant__ CUtensorMap tmVt, const FlashParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t QK_ATOM = 128 * 128;              // 128 rows x 128 B
    const uint32_t v_atom = (uint32_t)p.dn * 128u, v_stage = 2u * v_atom;
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + KA * QK_ATOM;               // 2 stages
    unsigned char* sV = sK + 2 * KA * QK_ATOM;           // 2 stages
    unsigned char* sP = sV + 2 * v_stage;                // 2 atoms of [128 x 64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * QK_ATOM);
    uint64_t* qfull = bars; uint64_t* kfull = bars + 1; uint64_t* vfull = bars + 3; uint64_t* mma_bar = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

    if (tid == 0) {
        for (int i = 0; i < 6; ++i) tc05::mbar_init(bars + i, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<256>(tmem_slot);
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *tmem_slot, tS = tmem, tO = tmem + 128u;
    const uint32_t trow = (uint32_t)(warp * 32) << 16;   // this thread's TMEM lane quarter
    const uint32_t sQ_s = tc05::smem_u32(sQ), sK_s = tc05::smem_u32(sK), sV_s = tc05::smem_u32(sV), sP_s = tc05::smem_u32(sP);
    const int T = (p.Lk + 127) / 128;
    const int ksteps = (p.d + 15) / 16;                  // K = 16 steps of the S MMA
    uint32_t kph[2] = {0u, 0u}, vph[2] = {0u, 0u}, mph = 0u;

    auto load_k = [&](int j, int st) {
        mbar_expect_tx(kfull + st, KA * QK_ATOM);
        for (int a = 0; a < KA; ++a) tma_load_4d(sK_s + (st * KA + a) * QK_ATOM, &tmK, kfull + st, a * 64, j * 128, h, b);
    };
    auto load_v = [&](int j, int st) {
        mbar_expect_tx(vfull + st, v_stage);
        for (int a = 0; a < 2; ++a) tma_load_4d(sV_s + st * v_stage + a * v_atom, &tmVt, vfull + st, j * 128 + a * 64, 0, h, b);
    };
    auto issue_s = [&](int st) {                          // S = Q K^T for the tile in stage st
        constexpr uint32_t idesc = tc05::idesc_f16(128, 128);
        for (int s = 0; s < ksteps; ++s) {
            const int a = s >> 2, k4 = s & 3;
            tc05::mma_f16(tS, smem_desc_sw128(sQ_s + a * QK_ATOM) + 2ull * k4, smem_desc_sw128(sK_s + (st * KA + a) * QK_ATOM) + 2ull * k4, idesc,
                          s != 0 ? 1u : 0u);
        }
        tc05::mma_commit(mma_bar);
    };

    if (tid == 0) {
        mbar_expect_tx(qfull, KA * QK_ATOM);
        for (int a = 0; a < KA; ++a) tma_load_4d(sQ_s + a * QK_ATOM, &tmQ, qfull, a * 64, q0, h, b);
        load_k(0, 0);
    }
    // ---- pass 1: row statistics ----
    const float s2 = p.scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        if (tid == 0) {
            if (j + 1 < T) load_k(j + 1, st ^ 1);        // that stage's last reader (S of tile j-1) completed before the last barrier
            if (j == 0) tc05::mbar_wait(qfull, 0u);
            tc05::mbar_wait(kfull + st, kph[st]); kph[st] ^= 1u;
            tc05::fence_after_sync();
            issue_s(st);
        }
        tc05::mbar_wait(mma_bar, mph); mph ^= 1u;
        __syncwarp();                                     // lane 0 of warp 0 rejoins before the warp-collective TMEM loads
        tc05::fence_after_sync();
        const int ncols = min(128, p.Lk - j * 128);
        // statistics in the log2 domain: exp(scale s - max) = 2^(s2 s - m), s2 = scale log2(e): one FFMA + one MUFU.EX2 per
        // score; the column mask only exists in the last (partial) key tile
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            float acc[16];
            tc05::tmem_ld16(tS + trow + q * 16, acc);
            if (q * 16 >= ncols) continue;
            if (q * 16 + 16 > ncols) {
#pragma unroll
                for (int c = 0; c < 16; ++c) if (q * 16 + c >= ncols) acc[c] = -INFINITY;
            }
            float cm = acc[0];
#pragma unroll
            for (int c = 1; c < 16; ++c) cm = fmaxf(cm, acc[c]);
            const float mn = fmaxf(m, cm * s2);              // scale > 0: the maximum commutes with the scaling
            float add = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c) add += ex2_approx(fmaf(acc[c], s2, -mn));
            l = fmaf(l, ex2_approx(m - mn), add);
            m = mn;
        }
        tc05::fence_before_sync();
        __syncthreads();                                  // everyone has read S before the next tile's MMA overwrites it
    }
    const float inv_l = 1.0f / l;
    // ---- pass 2: O = P V ----
    if (tid == 0) { tc05::fence_after_sync(); load_k(0, 0); load_v(0, 0); }
    const uint32_t idesc_o = tc05::idesc_f16(128, (uint32_t)p.dn);
    const int r = tid;                                    // query row of this thread = TMEM lane
    for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        if (tid == 0) {
            if (j + 1 < T) { load_k(j + 1, st ^ 1); load_v(j + 1, st ^ 1); }
            tc05::mbar_wait(kfull + st, kph[st]); kph[st] ^= 1u;
            tc05::fence_after_sync();
            issue_s(st);
        }
        tc05::mbar_wait(mma_bar, mph); mph ^= 1u;
        __syncwarp();                                     // lane 0 of warp 0 rejoins before the warp-collective TMEM loads
        tc05::fence_after_sync();
        const int ncols = min(128, p.Lk - j * 128);
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            float acc[16];
            tc05::tmem_ld16(tS + trow + q * 16, acc);
            float pr[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pr[c] = ex2_approx(fmaf(acc[c], s2, -m)) * inv_l;
            if (q * 16 + 16 > ncols) {
#pragma unroll
                for (int c = 0; c < 16; ++c) if (q * 16 + c >= ncols) pr[c] = 0.f;
            }
            // kv columns q*16 .. q*16+15 -> atom (q >> 2), 16-byte chunks 2*(q & 3), +1 of this row, XOR-swizzled by row % 8
            unsigned char* rowp = sP + (q >> 2) * QK_ATOM + (r >> 3) * 1024 + (r & 7) * 128;
            const int ch = 2 * (q & 3);
            uint4 lo, hi;
            lo.x = tc05::pack_f16x2(pr[0], pr[1]); lo.y = tc05::pack_f16x2(pr[2], pr[3]); lo.z = tc05::pack_f16x2(pr[4], pr[5]); lo.w = tc05::pack_f16x2(pr[6], pr[7]);
            hi.x = tc05::pack_f16x2(pr[8], pr[9]); hi.y = tc05::pack_f16x2(pr[10], pr[11]); hi.z = tc05::pack_f16x2(pr[12], pr[13]); hi.w = tc05::pack_f16x2(pr[14], pr[15]);
            *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) = lo;
            *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = hi;
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before_sync();
        __syncthreads();                                  // P complete and visible to the tensor core; S fully consumed
        if (tid == 0) {
            tc05::fence_after_sync();
            tc05::mbar_wait(vfull + st, vph[st]); vph[st] ^= 1u;
            tc05::fence_after_sync();
            for (int s = 0; s < 8; ++s) {                 // K = 128 keys = 8 x 16
                const int a = s >> 2, k4 = s & 3;
                tc05::mma_f16(tO, smem_desc_sw128(sP_s + a * QK_ATOM) + 2ull * k4, smem_desc_sw128(sV_s + st * v_stage + a * v_atom) + 2ull * k4, idesc_o,
                              (j | s) != 0 ? 1u : 0u);
            }
            tc05::mma_commit(mma_bar);
        }
        tc05::mbar_wait(mma_bar, mph); mph ^= 1u;         // P, this V stage and S are free again
        __syncwarp();
        tc05::fence_after_sync();
    }
    // ---- epilogue: O row -> fp16 (the TMEM loads are warp-collective: every lane executes them, only valid rows store) ----
    {
        const bool row_ok = q0 + r < p.Lq;
        __half* dst = p.out + ((long long)b * p.Lq + (row_ok ? q0 + r : 0)) * p.ld_out + (long long)h * p.d;
#pragma unroll 1
        for (int q = 0; q < p.dn / 16; ++q) {
            float acc[16];
            tc05::tmem_ld16(tO + trow + q * 16, acc);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int c = q * 16 + half * 8;
                if (row_ok && c + 8 <= p.d) {
                    uint4 o;
                    o.x = tc05::pack_f16x2(acc[half * 8 + 0], acc[half * 8 + 1]); o.y = tc05::pack_f16x2(acc[half * 8 + 2], acc[half * 8 + 3]);
                    o.z = tc05::pack_f16x2(acc[half * 8 + 4], acc[half * 8 + 5]); o.w = tc05::pack_f16x2(acc[half * 8 + 6], acc[half * 8 + 7]);
                    *reinterpret_cast<uint4*>(dst + c) = o;
                }
            }
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<256>(tmem);
}

inline bool make_map4(CUtensorMap* map, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3], cuuint32_t box0, cuuint32_t box1) {
    const cuuint32_t box[4] = {box0, box1, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return false;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------------------------------------
// Producers (all HBM-streaming, one element group per thread, 16-byte stores)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// GroupNorm statistics over NHWC: grid (chunks, B).  A thread owns channel c (+ blockDim strides) over a chunk of pixels,
// so a warp reads consecutive channels of one pixel; per-channel partials fold into their group in shared memory and
// leave through fp64 atomics: sums[b][g] = (sum, sum of squares).
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int HW, int C, int G, int px_per_block, double* __restrict__ sums) {
    extern __shared__ float shf[];                      // [C][2] per-channel partials of this block's pixels
    const int b = blockIdx.y, cpg = C / G;
    const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f, ss = 0.f;
        const float* col = x + ((long long)b * HW + p0) * C + c;
        for (int p = p0; p < p1; ++p, col += C) { const float v = *col; s += v; ss = fmaf(v, v, ss); }
        shf[2 * c] = s; shf[2 * c + 1] = ss;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s = 0.0, ss = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s += (double)shf[2 * c]; ss += (double)shf[2 * c + 1]; }
        atomicAdd(&sums[((long long)b * G + g) * 2], s);
        atomicAdd(&sums[((long long)b * G + g) * 2 + 1], ss);
    }
}

// Vectorised variant for C % 4 == 0, (C / G) % 4 == 0 and 256 % (C / 4) == 0 (every Stable-Diffusion shape): a thread owns FOUR
// consecutive channels (one 16-byte load per pixel, all in one group) and every (C / 4)-th ... pixel of the block's chunk, so
// all 256 threads are busy whatever C is; partials fold through shared memory into fp64 atomics per (batch, group).
__global__ void __launch_bounds__(256) gn_stats_vec_kernel(const float* __restrict__ x, int HW, int C, int G, int px_per_block, double* __restrict__ sums) {
    __shared__ float sh[256][2];
    const int q = C >> 2, cq = threadIdx.x % q, pr = threadIdx.x / q, rows = 256 / q;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
    float s = 0.f, ss = 0.f;
    for (int p = p0 + pr; p < p1; p += rows) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((long long)b * HW + p) * C) + cq);
        s += (v.x + v.y) + (v.z + v.w);
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    }
    sh[threadIdx.x][0] = s; sh[threadIdx.x][1] = ss;
    __syncthreads();
    // thread g < G sums the (cpg / 4) quads of its group over the `rows` pixel rows
    const int qpg = (C / G) >> 2;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double a = 0.0, a2 = 0.0;
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < qpg; ++k) { a += (double)sh[r * q + g * qpg + k][0]; a2 += (double)sh[r * q + g * qpg + k][1]; }
        atomicAdd(&sums[((long long)b * G + g) * 2], a);
        atomicAdd(&sums[((long long)b * G + g) * 2 + 1], a2);
    }
}

// (sum, sumsq) -> (mean, rstd), in place as floats: stats[b][g] = (mean, rstd).
__global__ void gn_finalize_kernel(const double* __restrict__ sums, int n, double count, float eps, float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mean = sums[2 * i] / count;
    double var = sums[2 * i + 1] / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

struct NormArgs {          // optional GroupNorm(+SiLU) applied while reading the source
    const float* stats;    // [B][G][2] (mean, rstd) or NULL
    const float* gamma; const float* beta;
    int G, act;
};

__device__ __forceinline__ void norm8(float (&v)[8], const NormArgs& nm, int b, int c0, int C) {
    if (!nm.stats) return;
    const int cpg = C / nm.G;
    if ((cpg & 3) == 0) {           // channels c0..c0+3 and c0+4..c0+7 each lie in one group: two statistics, 16-byte parameter loads
        const float* st = nm.stats + (long long)b * nm.G * 2;
        const int g0 = c0 / cpg, g1 = (c0 + 4) / cpg;
        const float m0 = st[2 * g0], r0 = st[2 * g0 + 1], m1 = st[2 * g1], r1 = st[2 * g1 + 1];
        const float4 ga0 = __ldg(reinterpret_cast<const float4*>(nm.gamma + c0)), ga1 = __ldg(reinterpret_cast<const float4*>(nm.gamma + c0 + 4));
        const float4 be0 = __ldg(reinterpret_cast<const float4*>(nm.beta + c0)), be1 = __ldg(reinterpret_cast<const float4*>(nm.beta + c0 + 4));
        const float ga[8] = {ga0.x, ga0.y, ga0.z, ga0.w, ga1.x, ga1.y, ga1.z, ga1.w};
        const float be[8] = {be0.x, be0.y, be0.z, be0.w, be1.x, be1.y, be1.z, be1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float y = (v[j] - (j < 4 ? m0 : m1)) * (j < 4 ? r0 : r1) * ga[j] + be[j];
            v[j] = nm.act ? silu(y) : y;
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j, g = c / cpg;
        const float mean = nm.stats[((long long)b * nm.G + g) * 2], rstd = nm.stats[((long long)b * nm.G + g) * 2 + 1];
        float y = (v[j] - mean) * rstd * nm.gamma[c] + nm.beta[c];
        v[j] = nm.act ? silu(y) : y;
    }
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint4 o;
    o.x = tc05::pack_f16x2(v[0], v[1]); o.y = tc05::pack_f16x2(v[2], v[3]);
    o.z = tc05::pack_f16x2(v[4], v[5]); o.w = tc05::pack_f16x2(v[6], v[7]);
    return o;
}

// NHWC fp32 [B,Hs,Ws,C] -> fp16 [B*Ho*Wo, ks*ks*Cp] (K index = (ky*ks + kx)*Cp + c; Cp = C rounded up to 8, zero padded),
// zero padding `pad` on the top/left (bottom/right implied by Ho, Wo), stride 1/2, optional nearest x2 up-sampling of the
// source (Upsample2D), optional GroupNorm(+SiLU) on the fly.  ks = 1 is the plain normalise-and-cast producer.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, int B, int Hs, int Ws, int C, int Cp, int ks, int stride, int pad,
                                                     int up, int Ho, int Wo, const NormArgs nm, __half* __restrict__ out) {
    const int chunks = Cp / 8;
    const long long total = (long long)B * Ho * Wo * ks * ks * chunks;
    const int Hin = up ? Hs * 2 : Hs, Win = up ? Ws * 2 : Ws;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(t % chunks);
        long long r = t / chunks;
        const int tap = (int)(r % (ks * ks)); r /= ks * ks;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const int b = (int)(r / Ho);
        const int iy = oy * stride + tap / ks - pad, ix = ox * stride + tap % ks - pad;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
            const int sy = up ? iy >> 1 : iy, sx = up ? ix >> 1 : ix;
            const float* src = x + (((long long)b * Hs + sy) * Ws + sx) * C + ch * 8;
            if (ch * 8 + 8 <= C && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                const float4 a = reinterpret_cast<const float4*>(src)[0], c4 = reinterpret_cast<const float4*>(src)[1];
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
                norm8(v, nm, b, ch * 8, C);
            } else {
                for (int j = 0; j < 8; ++j) {
                    if (ch * 8 + j < C) {
                        float one[8] = {src[j], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (nm.stats) {
                            const int c = ch * 8 + j, g = c / (C / nm.G);
                            const float y = (one[0] - nm.stats[((long long)b * nm.G + g) * 2]) * nm.stats[((long long)b * nm.G + g) * 2 + 1] * nm.gamma[c] + nm.beta[c];
                            one[0] = nm.act ? silu(y) : y;
                        }
                        v[j] = one[0];
                    }
                }
            }
        }
        reinterpret_cast<uint4*>(out)[t] = pack8(v);
    }
}

// The ks = 1 case of im2col_kernel on its own (the producer of every implicit-GEMM convolution and projection: GroupNorm(+SiLU)
// on the fly, fp32 NHWC -> fp16 NHWC) for C % 8 == 0: one 8-channel chunk per thread, no tap / stride / padding index math.
__global__ void __launch_bounds__(256) norm_cast_kernel(const float* __restrict__ x, long long total_chunks, int HW, int C, const NormArgs nm,
                                                        __half* __restrict__ out) {
    const int chunks = C >> 3;
    const long long per_image = (long long)HW * chunks;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total_chunks; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / per_image);
        const int ch = (int)((t - (long long)b * per_image) % chunks);
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * t), c4 = __ldg(reinterpret_cast<const float4*>(x) + 2 * t + 1);
        float v[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
        norm8(v, nm, b, ch * 8, C);
        reinterpret_cast<uint4*>(out)[t] = pack8(v);
    }
}

// LayerNorm over the last axis, fp32 [M,C] -> fp16 [M,C]; one warp per row, two passes over registers/L1.
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, int M, int C, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, __half* __restrict__ out) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + (long long)row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; ss = fmaf(d, d, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)C + eps);
    for (int c = lane; c < C; c += 32) out[(long long)row * C + c] = __float2half_rn((xr[c] - mean) * rstd * gamma[c] + beta[c]);
}

// GEGLU: x [M, 2I] fp32 -> fp16 [M, I] = value * gelu(gate), exact (erf) GELU like F.gelu.
__global__ void __launch_bounds__(256) geglu_kernel(const float* __restrict__ x, long long M, int I, __half* __restrict__ out) {
    const long long total = M * I;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long m = t / I;
        const int i = (int)(t - m * I);
        const float v = x[m * 2 * I + i], g = x[m * 2 * I + I + i];
        out[t] = __float2half_rn(v * (0.5f * g * (1.0f + erff(g * 0.70710678118654752f))));
    }
}

// Row softmax of scale * scores: fp32 [rows, L] (row stride ld_in) -> fp16 [rows, ld_out] (columns >= L zeroed).
// One warp per row; the row is read once from HBM/L2 (16-byte loads, 512 B per warp instruction, when the row is
// 16 B aligned) and twice more from L1.
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ s, long long rows, int L, long long ld_in, long long ld_out, float scale,
                                                      __half* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* sr = s + row * ld_in;
    __half* orow = out + row * ld_out;
    const bool vec = (L & 3) == 0 && (ld_in & 3) == 0 && (ld_out & 3) == 0 && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    float mx = -INFINITY, sum = 0.f;
    if (vec) {
        const float4* s4 = reinterpret_cast<const float4*>(sr);
        const int L4 = L >> 2;
#pragma unroll 4
        for (int c = lane; c < L4; c += 32) { const float4 v = s4[c]; mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w))); }
        mx *= scale;                                     // scale > 0: max commutes with the scaling
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll 4
        for (int c = lane; c < L4; c += 32) {
            const float4 v = s4[c];
            sum += __expf(v.x * scale - mx) + __expf(v.y * scale - mx) + __expf(v.z * scale - mx) + __expf(v.w * scale - mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        uint2* o2 = reinterpret_cast<uint2*>(orow);
#pragma unroll 4
        for (int c = lane; c < L4; c += 32) {
            const float4 v = s4[c];
            uint2 p;
            p.x = tc05::pack_f16x2(__expf(v.x * scale - mx) * inv, __expf(v.y * scale - mx) * inv);
            p.y = tc05::pack_f16x2(__expf(v.z * scale - mx) * inv, __expf(v.w * scale - mx) * inv);
            o2[c] = p;
        }
        for (int c = L + lane; c < (int)ld_out; c += 32) orow[c] = __half(0.f);
        return;
    }
    for (int c = lane; c < L; c += 32) mx = fmaxf(mx, sr[c] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    for (int c = lane; c < L; c += 32) sum += __expf(sr[c] * scale - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int c = lane; c < (int)ld_out; c += 32) orow[c] = c < L ? __float2half_rn(__expf(sr[c] * scale - mx) * inv) : __half(0.f);
}

__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ x, long long n, __half* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) out[t] = __float2half_rn(x[t]);
}

// ---- backward of GroupNorm(+SiLU) over NHWC (the VAE encoder's input gradient IS the SDS gradient) ----------------------
// y = act(gamma * xhat + beta), xhat = (x - mean) rstd.  With dz = dy * act'(z):
//   dx = rstd * (gamma dz - s1 / n - xhat s2 / n),  s1 = sum_group gamma dz,  s2 = sum_group gamma dz xhat,  n = HW * C / G.
__device__ __forceinline__ float gn_dz(float x, float dy, float mean, float rstd, float g, float b, int act, float& xhat) {
    xhat = (x - mean) * rstd;
    if (!act) return dy;
    const float z = fmaf(g, xhat, b), sg = 1.0f / (1.0f + __expf(-z));
    return dy * (sg * (1.0f + z * (1.0f - sg)));
}
// pass 1: grid (chunks, B), same mapping as gn_stats_kernel; sums[b][g] = (s1, s2) through fp64 atomics
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int act, int HW, int C, int G,
                                                            int px_per_block, double* __restrict__ sums) {
    extern __shared__ float shf[];
    const int b = blockIdx.y, cpg = C / G;
    const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1], ga = gamma[c], be = beta[c];
        float s1 = 0.f, s2 = 0.f;
        long long at = ((long long)b * HW + p0) * C + c;
        for (int p = p0; p < p1; ++p, at += C) {
            float xh;
            const float dz = gn_dz(x[at], dy[at], mean, rstd, ga, be, act, xh) * ga;
            s1 += dz; s2 = fmaf(dz, xh, s2);
        }
        shf[2 * c] = s1; shf[2 * c + 1] = s2;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s1 += (double)shf[2 * c]; s2 += (double)shf[2 * c + 1]; }
        atomicAdd(&sums[((long long)b * G + g) * 2], s1);
        atomicAdd(&sums[((long long)b * G + g) * 2 + 1], s2);
    }
}
// pass 2: elementwise; dx (+ add) -> fp32 and / or fp16
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, int act, long long total,
                                                           int HW, int C, int G, const double* __restrict__ sums, const float* __restrict__ add,
                                                           float* __restrict__ dx32, __half* __restrict__ dx16) {
    const int cpg = C / G;
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const int b = (int)(t / ((long long)HW * C)), g = c / cpg;
        const long long sg = ((long long)b * G + g) * 2;
        const float mean = stats[sg], rstd = stats[sg + 1], ga = gamma[c];
        float xh;
        const float dz = gn_dz(x[t], dy[t], mean, rstd, ga, beta[c], act, xh) * ga;
        float v = rstd * (dz - (float)sums[sg] * inv_n - xh * (float)sums[sg + 1] * inv_n);
        if (add) v += add[t];
        if (dx32) dx32[t] = v;
        if (dx16) dx16[t] = __float2half_rn(v);
    }
}

// Vectorised variants (same shape conditions as gn_stats_vec_kernel): four channels of one group per thread, 16-byte accesses.
__global__ void __launch_bounds__(256) gn_bwd_reduce_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta, int act, int HW, int C,
                                                                int G, int px_per_block, double* __restrict__ sums) {
    __shared__ float sh[256][2];
    const int q = C >> 2, cq = threadIdx.x % q, pr = threadIdx.x / q, rows = 256 / q;
    const int b = blockIdx.y, g = (4 * cq) / (C / G);
    const int p0 = blockIdx.x * px_per_block, p1 = min(p0 + px_per_block, HW);
    const float mean = stats[((long long)b * G + g) * 2], rstd = stats[((long long)b * G + g) * 2 + 1];
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cq), be = __ldg(reinterpret_cast<const float4*>(beta) + cq);
    float s1 = 0.f, s2 = 0.f;
    for (int p = p0 + pr; p < p1; p += rows) {
        const long long at = ((long long)b * HW + p) * (long long)q + cq;
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + at), dv = __ldg(reinterpret_cast<const float4*>(dy) + at);
        float xh, dz;
        dz = gn_dz(xv.x, dv.x, mean, rstd, ga.x, be.x, act, xh) * ga.x; s1 += dz; s2 = fmaf(dz, xh, s2);
        dz = gn_dz(xv.y, dv.y, mean, rstd, ga.y, be.y, act, xh) * ga.y; s1 += dz; s2 = fmaf(dz, xh, s2);
        dz = gn_dz(xv.z, dv.z, mean, rstd, ga.z, be.z, act, xh) * ga.z; s1 += dz; s2 = fmaf(dz, xh, s2);
        dz = gn_dz(xv.w, dv.w, mean, rstd, ga.w, be.w, act, xh) * ga.w; s1 += dz; s2 = fmaf(dz, xh, s2);
    }
    sh[threadIdx.x][0] = s1; sh[threadIdx.x][1] = s2;
    __syncthreads();
    const int qpg = (C / G) >> 2;
    for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
        double a = 0.0, a2 = 0.0;
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < qpg; ++k) { a += (double)sh[r * q + gg * qpg + k][0]; a2 += (double)sh[r * q + gg * qpg + k][1]; }
        atomicAdd(&sums[((long long)b * G + gg) * 2], a);
        atomicAdd(&sums[((long long)b * G + gg) * 2 + 1], a2);
    }
}
__global__ void __launch_bounds__(256) gn_bwd_apply_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta, int act, long long total4,
                                                               int HW, int C, int G, const double* __restrict__ sums, const float* __restrict__ add,
                                                               float* __restrict__ dx32, __half* __restrict__ dx16) {
    const int q = C >> 2, cpg = C / G;
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    const long long per_image = (long long)HW * q;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / per_image);
        const int cq = (int)((t - (long long)b * per_image) % q);
        const long long sg = ((long long)b * G + (4 * cq) / cpg) * 2;
        const float mean = stats[sg], rstd = stats[sg + 1];
        const float m1 = (float)sums[sg] * inv_n, m2 = (float)sums[sg + 1] * inv_n;
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cq), be = __ldg(reinterpret_cast<const float4*>(beta) + cq);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + t), dv = __ldg(reinterpret_cast<const float4*>(dy) + t);
        float4 o;
        float xh, dz;
        dz = gn_dz(xv.x, dv.x, mean, rstd, ga.x, be.x, act, xh) * ga.x; o.x = rstd * (dz - m1 - xh * m2);
        dz = gn_dz(xv.y, dv.y, mean, rstd, ga.y, be.y, act, xh) * ga.y; o.y = rstd * (dz - m1 - xh * m2);
        dz = gn_dz(xv.z, dv.z, mean, rstd, ga.z, be.z, act, xh) * ga.z; o.z = rstd * (dz - m1 - xh * m2);
        dz = gn_dz(xv.w, dv.w, mean, rstd, ga.w, be.w, act, xh) * ga.w; o.w = rstd * (dz - m1 - xh * m2);
        if (add) { const float4 av = __ldg(reinterpret_cast<const float4*>(add) + t); o.x += av.x; o.y += av.y; o.z += av.z; o.w += av.w; }
        if (dx32) reinterpret_cast<float4*>(dx32)[t] = o;
        if (dx16) { uint2 h; h.x = tc05::pack_f16x2(o.x, o.y); h.y = tc05::pack_f16x2(o.z, o.w); reinterpret_cast<uint2*>(dx16)[t] = h; }
    }
}

// softmax backward: dS = P (dP - sum_row(P dP)) * scale.  One warp per row; P fp16 [rows, ld], dP fp32 [rows, ld] -> dS fp16.
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __half* __restrict__ P, const float* __restrict__ dP, long long rows, int L, long long ld,
                                                          float scale, __half* __restrict__ dS) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const __half* pr = P + row * ld;
    const float* dr = dP + row * ld;
    float r = 0.f;
    for (int c = lane; c < L; c += 32) r = fmaf(__half2float(pr[c]), dr[c], r);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    for (int c = lane; c < (int)ld; c += 32) dS[row * ld + c] = c < L ? __float2half_rn(__half2float(pr[c]) * (dr[c] - r) * scale) : __half(0.f);
}

// fp16 [rows, cols] (row stride ld_in) -> [cols, rows] (row stride ld_out), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) transpose_f16_kernel(const __half* __restrict__ in, int rows, int cols, long long ld_in, __half* __restrict__ out,
                                                            long long ld_out) {
    __shared__ __half tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(long long)(r0 + j) * ld_in + c0 + tx];
    __syncthreads();
    for (int j = ty; j < 32; j += 8)
        if (c0 + j < cols && r0 + tx < rows) out[(long long)(c0 + j) * ld_out + r0 + tx] = tile[tx][j];
}

// Operand of the input gradient of a 3x3 / stride-2 convolution with zero padding on the bottom / right only (the VAE's
// Downsample2D): dy fp32 NHWC [B,Ho,Wo,N] -> fp16 [B*H*W, 9*Np]; row (b, y, x), K index (ky*3 + kx)*Np + n holds
// dy[b, (y-ky)/2, (x-kx)/2, n] when y-ky and x-kx are even, non-negative and inside, else 0.  A GEMM with
// W[c][(ky*3+kx)*Np + n] = w[n][c][ky][kx] then gives d_in [B*H*W, C].
__global__ void __launch_bounds__(256) conv_s2_dgrad_operand_kernel(const float* __restrict__ dy, int B, int Ho, int Wo, int N, int Np, int H, int W,
                                                                    __half* __restrict__ out) {
    const int chunks = Np / 8;
    const long long total = (long long)B * H * W * 9 * chunks;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(t % chunks);
        long long r = t / chunks;
        const int tap = (int)(r % 9); r /= 9;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int b = (int)(r / H);
        const int yy = y - tap / 3, xx = x - tap % 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (yy >= 0 && xx >= 0 && !(yy & 1) && !(xx & 1) && (yy >> 1) < Ho && (xx >> 1) < Wo) {
            const float* src = dy + (((long long)b * Ho + (yy >> 1)) * Wo + (xx >> 1)) * N + ch * 8;
            for (int j = 0; j < 8; ++j) if (ch * 8 + j < N) v[j] = src[j];
        }
        reinterpret_cast<uint4*>(out)[t] = pack8(v);
    }
}

// shapes the vectorised GroupNorm kernels take: four channels per thread inside one group, C / 4 lanes dividing the block
inline bool gn_vec_ok(int C, int G) { return C % 4 == 0 && (C / G) % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0; }

// pixels per block of the vectorised GroupNorm reductions: ~512 blocks per image whatever the resolution -- enough to fill the
// machine, few enough that the fp64 atomics onto the 2 G group sums (one pair per block and group) do not queue up
inline int gn_vec_pixels(int HW) { const int p = HW / 512; return p < 16 ? 16 : (p > 512 ? 512 : p); }

inline int grid_for(long long work, int block, int per_sm) {
    const long long want = (work + block - 1) / block;
    const long long cap = (long long)acb::sm_count() * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" {

int ac_sd_gemm_f16(const void* A, const void* W, const float* bias, const float* group_bias, int rows_per_group, const float* residual,
                   void* C, int out_f16, int M, int N, int K, int64_t lda, int64_t ldw, int64_t ldc, int64_t ldr, int batch_outer,
                   int batch_inner, int64_t sAo, int64_t sAi, int64_t sWo, int64_t sWi, int64_t sCo, int64_t sCi, void* stream) {
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || batch_outer <= 0 || batch_inner <= 0) return AC_E_INVALID_ARG;
    if ((lda & 7) || (ldw & 7) || (sAo & 7) || (sAi & 7) || (sWo & 7) || (sWi & 7)) return AC_E_INVALID_ARG;     // 16 B operand rows
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return AC_E_INVALID_ARG;
    if (group_bias && rows_per_group <= 0) return AC_E_INVALID_ARG;
    if ((batch_outer > 1 || batch_inner > 1) && (residual || group_bias)) return AC_E_UNSUPPORTED;
    ACB_SET_MAX_SMEM(sd_gemm_kernel, GEMM_SMEM);
    GemmParams p;
    p.A = reinterpret_cast<const __half*>(A); p.W = reinterpret_cast<const __half*>(W);
    p.bias = bias; p.group_bias = group_bias; p.residual = residual; p.C = C;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw; p.ldc = ldc; p.ldr = ldr;
    p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; p.batch_inner = batch_inner;
    p.sAo = sAo; p.sAi = sAi; p.sWo = sWo; p.sWi = sWi; p.sCo = sCo; p.sCi = sCi; p.out_f16 = out_f16;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch_outer * batch_inner);
    // Split-K when the output tiles alone cannot fill the machine (weight-streaming GEMMs: a handful of tiles, K in the
    // thousands): about two CTAs per SM, at least 4 k-tiles each; fp32 outputs only (partials meet in atomics).
    p.splits = 1;
    const long long tiles = (long long)grid.x * grid.y, KT_all = (K + BK - 1) / BK;
    if (grid.z == 1 && !out_f16 && tiles < acb::sm_count() && KT_all >= 8) {
        long long want = (2LL * acb::sm_count() + tiles - 1) / tiles;
        if (want > KT_all / 4) want = KT_all / 4;
        if (want > 1) {
            p.splits = (int)want;
            grid.z = (unsigned)want;
            if (ldc == N) { if (cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, (cudaStream_t)stream) != cudaSuccess) return acb::cuda_fail(); }
            else { if (cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, (cudaStream_t)stream) != cudaSuccess) return acb::cuda_fail(); }
        }
    }
    if (grid.y > 65535 || grid.z > 65535) return AC_E_INVALID_ARG;
    static const bool use_cp_async = [] { const char* e = getenv("AC_SD_GEMM"); return e && e[0] == 'c'; }();   // A/B switch: AC_SD_GEMM=cpasync
    if (use_cp_async) {
        sd_gemm_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(p);
        return acb::launched();
    }
    ACB_SET_MAX_SMEM(sd_gemm_tma_kernel, GEMM_SMEM);
    CUtensorMap tmA, tmW;
    int a_zi, a_zo, w_zi, w_zo;
    if (!make_operand_map(&tmA, A, K, M, lda, batch_inner, sAi, batch_outer, sAo, &a_zi, &a_zo) ||
        !make_operand_map(&tmW, W, K, N, ldw, batch_inner, sWi, batch_outer, sWo, &w_zi, &w_zo))
        return AC_E_UNSUPPORTED;                         // no silent fallback: the caller sees the failure
    sd_gemm_tma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(tmA, tmW, p, a_zi, a_zo, w_zi, w_zo, ConvGeom{0, 0, 0, 0});
    return acb::launched();
}

int ac_sd_conv3x3_f16(const void* act, const void* W, const float* bias, const float* group_bias, const float* residual, float* out, int B, int H,
                      int Wd, int C, int N, void* stream) {
    // 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM: act fp16 NHWC [B,H,Wd,C], W fp16 [N, 9*C] (tap-major),
    // out fp32 NHWC [B,H,Wd,N] (+ bias[N], + group_bias[B,N], + residual [B,H,Wd,N]).
    if (!act || !W || !out || B <= 0 || H <= 0 || Wd <= 0 || C <= 0 || N <= 0 || (C & 63)) return AC_E_INVALID_ARG;
    if (((uintptr_t)act | (uintptr_t)W) & 15) return AC_E_INVALID_ARG;
    // the 128 output pixels of a tile must be whole image rows, or 128 consecutive pixels of one row when rows are wider
    // (the VAE's 256 / 512-pixel maps): (bw, bh, bb) = box extents over (W, H, B)
    int bw = Wd, bh, bb = 1;
    if (Wd > 128) {
        if (Wd % 128) return AC_E_UNSUPPORTED;
        bw = 128; bh = 1;
    } else {
        if (128 % Wd) return AC_E_UNSUPPORTED;
        bh = 128 / Wd;
        if (bh > H) { if (bh % H) return AC_E_UNSUPPORTED; bb = bh / H; bh = H; }
        else if (H % bh) return AC_E_UNSUPPORTED;
    }
    const long long M = (long long)B * H * Wd;
    if (M > 0x7FFFFFFF) return AC_E_INVALID_ARG;
    GemmParams p;
    p.A = nullptr; p.W = reinterpret_cast<const __half*>(W);
    p.bias = bias; p.group_bias = group_bias; p.residual = residual; p.C = out;
    p.M = (int)M; p.N = N; p.K = 9 * C; p.lda = 0; p.ldw = 9LL * C; p.ldc = N; p.ldr = N;
    p.rows_per_group = H * Wd; p.batch_inner = 1;
    p.sAo = p.sAi = p.sWo = p.sWi = p.sCo = p.sCi = 0; p.out_f16 = 0; p.splits = 1;
    dim3 grid((N + BN - 1) / BN, (unsigned)((M + BM - 1) / BM), 1);
    const long long tiles = (long long)grid.x * grid.y, KT_all = 9LL * C / BK;
    if (tiles < acb::sm_count() && KT_all >= 8) {
        long long want = (2LL * acb::sm_count() + tiles - 1) / tiles;
        if (want > KT_all / 4) want = KT_all / 4;
        if (want > 1) {
            p.splits = (int)want; grid.z = (unsigned)want;
            if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M * N, (cudaStream_t)stream) != cudaSuccess) return acb::cuda_fail();
        }
    }
    if (grid.y > 65535) return AC_E_INVALID_ARG;
    CUtensorMap tmA, tmW;
    {
        const cuuint64_t da[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t sa[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wd * C * 2, (cuuint64_t)H * Wd * C * 2};
        const cuuint32_t box[4] = {64u, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
        const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
        EncodeTiledFn fn = encode_tiled();
        if (!fn || fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(act), da, sa, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return AC_E_UNSUPPORTED;
        int zi, zo;
        if (!make_operand_map(&tmW, W, 9 * C, N, 9LL * C, 1, 0, 1, 0, &zi, &zo)) return AC_E_UNSUPPORTED;
    }
    ACB_SET_MAX_SMEM(sd_gemm_tma_kernel, GEMM_SMEM);
    sd_gemm_tma_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, (cudaStream_t)stream>>>(tmA, tmW, p, 0, 0, 0, 0, ConvGeom{9, C / 64, Wd, H});
    return acb::launched();
}

int ac_sd_flash_attention_f16(const void* q, const void* k, const void* vt, void* out, int B, int heads, int Lq, int Lk, int d, int64_t ld_q,
                              int64_t ld_k, int64_t ld_vt, int64_t ld_out, float scale, void* stream) {
    if (!q || !k || !vt || !out || B <= 0 || heads <= 0 || Lq <= 0 || Lk <= 0 || d <= 0 || (d & 7) || d > 128 || !(scale > 0.f)) return AC_E_INVALID_ARG;
    if ((ld_q & 7) || (ld_k & 7) || (ld_vt & 7) || (ld_out & 7) || ld_vt < Lk || ld_q < (int64_t)heads * d || ld_k < (int64_t)heads * d) return AC_E_INVALID_ARG;
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)vt | (uintptr_t)out) & 15) return AC_E_INVALID_ARG;
    if (heads > 65535 || B > 65535) return AC_E_INVALID_ARG;
    const int dn = (d + 15) / 16 * 16, KA = d <= 64 ? 1 : 2;
    CUtensorMap tmQ, tmK, tmV;
    {
        const cuuint64_t dq[4] = {(cuuint64_t)d, (cuuint64_t)Lq, (cuuint64_t)heads, (cuuint64_t)B};
        const cuuint64_t sq[3] = {(cuuint64_t)ld_q * 2, (cuuint64_t)d * 2, (cuuint64_t)Lq * ld_q * 2};
        const cuuint64_t dk[4] = {(cuuint64_t)d, (cuuint64_t)Lk, (cuuint64_t)heads, (cuuint64_t)B};
        const cuuint64_t sk[3] = {(cuuint64_t)ld_k * 2, (cuuint64_t)d * 2, (cuuint64_t)Lk * ld_k * 2};
        const cuuint64_t dv[4] = {(cuuint64_t)Lk, (cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)B};
        const cuuint64_t sv[3] = {(cuuint64_t)ld_vt * 2, (cuuint64_t)d * ld_vt * 2, (cuuint64_t)heads * d * ld_vt * 2};
        if (!make_map4(&tmQ, q, dq, sq, 64, 128) || !make_map4(&tmK, k, dk, sk, 64, 128) || !make_map4(&tmV, vt, dv, sv, 64, (cuuint32_t)dn))
            return AC_E_UNSUPPORTED;
    }
    FlashParams p;
    p.out = reinterpret_cast<__half*>(out); p.ld_out = ld_out; p.Lq = Lq; p.Lk = Lk; p.d = d; p.dn = dn; p.scale = scale;
    const size_t smem = (size_t)KA * 16384 * 3 + 4 * (size_t)dn * 128 + 2 * 16384 + 64;
    dim3 grid((Lq + 127) / 128, heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (KA == 1) {
        ACB_SET_MAX_SMEM(sd_flash_attn_kernel<1>, 227 * 1024);
        sd_flash_attn_kernel<1><<<grid, 128, smem, st>>>(tmQ, tmK, tmV, p);
    } else {
        ACB_SET_MAX_SMEM(sd_flash_attn_kernel<2>, 227 * 1024);
        sd_flash_attn_kernel<2><<<grid, 128, smem, st>>>(tmQ, tmK, tmV, p);
    }
    return acb::launched();
}

int ac_sd_group_norm_stats(const float* x, int B, int HW, int C, int G, float eps, double* sums_workspace, float* stats, void* stream) {
    if (!x || !sums_workspace || !stats || B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G) return AC_E_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(sums_workspace, 0, sizeof(double) * 2 * B * G, st) != cudaSuccess) return acb::cuda_fail();
    const int px = HW >= 2048 ? 32 : (HW >= 256 ? 8 : 2);     // enough blocks to fill the machine at every resolution
    dim3 grid((HW + px - 1) / px, B);
    if (gn_vec_ok(C, G) && ((uintptr_t)x & 15) == 0) {
        const int pxv = gn_vec_pixels(HW);
        gn_stats_vec_kernel<<<dim3((HW + pxv - 1) / pxv, B), 256, 0, st>>>(x, HW, C, G, pxv, sums_workspace);
    } else {
        gn_stats_kernel<<<grid, 256, sizeof(float) * 2 * C, st>>>(x, HW, C, G, px, sums_workspace);
    }
    int rc = acb::launched();
    if (rc) return rc;
    gn_finalize_kernel<<<(B * G + 127) / 128, 128, 0, st>>>(sums_workspace, B * G, (double)HW * (C / G), eps, stats);
    return acb::launched();
}

int ac_sd_im2col_f16(const float* x, int B, int Hs, int Ws, int C, int ksize, int stride, int pad, int upsample2x, int Ho, int Wo,
                     const float* gn_stats, const float* gn_gamma, const float* gn_beta, int gn_groups, int gn_silu, void* out, void* stream) {
    if (!x || !out || B <= 0 || Hs <= 0 || Ws <= 0 || C <= 0 || Ho <= 0 || Wo <= 0 || (ksize != 1 && ksize != 3) || (stride != 1 && stride != 2))
        return AC_E_INVALID_ARG;
    if (gn_stats && (!gn_gamma || !gn_beta || gn_groups <= 0 || C % gn_groups)) return AC_E_INVALID_ARG;
    NormArgs nm; nm.stats = gn_stats; nm.gamma = gn_gamma; nm.beta = gn_beta; nm.G = gn_groups > 0 ? gn_groups : 1; nm.act = gn_silu;
    const int Cp = (C + 7) / 8 * 8;
    const long long total = (long long)B * Ho * Wo * ksize * ksize * (Cp / 8);
    if (ksize == 1 && stride == 1 && pad == 0 && !upsample2x && Ho == Hs && Wo == Ws && (C & 7) == 0 && ((uintptr_t)x & 15) == 0 &&
        (!gn_stats || (((uintptr_t)gn_gamma | (uintptr_t)gn_beta) & 15) == 0)) {
        norm_cast_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, total, Hs * Ws, C, nm, reinterpret_cast<__half*>(out));
        return acb::launched();
    }
    im2col_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, B, Hs, Ws, C, Cp, ksize, stride, pad, upsample2x, Ho, Wo, nm,
                                                                             reinterpret_cast<__half*>(out));
    return acb::launched();
}

int ac_sd_layer_norm_f16(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* out, void* stream) {
    if (!x || !gamma || !beta || !out || M <= 0 || C <= 0) return AC_E_INVALID_ARG;
    layer_norm_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, M, C, gamma, beta, eps, reinterpret_cast<__half*>(out));
    return acb::launched();
}

int ac_sd_geglu_f16(const float* x, int64_t M, int inner, void* out, void* stream) {
    if (!x || !out || M <= 0 || inner <= 0) return AC_E_INVALID_ARG;
    geglu_kernel<<<grid_for(M * inner, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, M, inner, reinterpret_cast<__half*>(out));
    return acb::launched();
}

int ac_sd_softmax_f16(const float* scores, int64_t rows, int L, int64_t ld_in, int64_t ld_out, float scale, void* out, void* stream) {
    if (!scores || !out || rows <= 0 || L <= 0 || ld_in < L || ld_out < L || !(scale > 0.0f)) return AC_E_INVALID_ARG;
    const long long blocks = (rows + 7) / 8;
    if (blocks > 0x7FFFFFFFll) return AC_E_INVALID_ARG;
    softmax_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(scores, rows, L, ld_in, ld_out, scale, reinterpret_cast<__half*>(out));
    return acb::launched();
}

int ac_sd_cast_f16(const float* x, int64_t n, void* out, void* stream) {
    if (!x || !out || n <= 0) return AC_E_INVALID_ARG;
    cast_f16_kernel<<<grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<__half*>(out));
    return acb::launched();
}

int ac_sd_group_norm_backward(const float* x, const float* dy, int B, int HW, int C, int G, const float* stats, const float* gamma, const float* beta,
                              int silu_act, const float* add, float* dx32, void* dx16, double* sums_workspace, void* stream) {
    if (!x || !dy || !stats || !gamma || !beta || !sums_workspace || (!dx32 && !dx16) || B <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % G)
        return AC_E_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(sums_workspace, 0, sizeof(double) * 2 * B * G, st) != cudaSuccess) return acb::cuda_fail();
    const long long total = (long long)B * HW * C;
    const uintptr_t al = (uintptr_t)x | (uintptr_t)dy | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)add | (uintptr_t)dx32 | ((uintptr_t)dx16 << 1);
    if (gn_vec_ok(C, G) && (al & 15) == 0) {
        const int pxv = gn_vec_pixels(HW);
        gn_bwd_reduce_vec_kernel<<<dim3((HW + pxv - 1) / pxv, B), 256, 0, st>>>(x, dy, stats, gamma, beta, silu_act, HW, C, G, pxv, sums_workspace);
        if (int rc = acb::launched()) return rc;
        gn_bwd_apply_vec_kernel<<<grid_for(total / 4, 256, 16), 256, 0, st>>>(x, dy, stats, gamma, beta, silu_act, total / 4, HW, C, G, sums_workspace,
                                                                               add, dx32, reinterpret_cast<__half*>(dx16));
        return acb::launched();
    }
    const int px = HW >= 2048 ? 32 : (HW >= 256 ? 8 : 2);
    dim3 grid((HW + px - 1) / px, B);
    gn_bwd_reduce_kernel<<<grid, 256, sizeof(float) * 2 * C, st>>>(x, dy, stats, gamma, beta, silu_act, HW, C, G, px, sums_workspace);
    if (int rc = acb::launched()) return rc;
    gn_bwd_apply_kernel<<<grid_for(total, 256, 16), 256, 0, st>>>(x, dy, stats, gamma, beta, silu_act, total, HW, C, G, sums_workspace, add, dx32,
                                                                   reinterpret_cast<__half*>(dx16));
    return acb::launched();
}

int ac_sd_softmax_backward_f16(const void* probs, const float* dprobs, int64_t rows, int L, int64_t ld, float scale, void* dscores, void* stream) {
    if (!probs || !dprobs || !dscores || rows <= 0 || L <= 0 || ld < L) return AC_E_INVALID_ARG;
    const long long blocks = (rows + 7) / 8;
    if (blocks > 0x7FFFFFFFll) return AC_E_INVALID_ARG;
    softmax_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(probs), dprobs, rows, L, ld, scale,
                                                                             reinterpret_cast<__half*>(dscores));
    return acb::launched();
}

int ac_sd_transpose_f16(const void* in, int rows, int cols, int64_t ld_in, void* out, int64_t ld_out, void* stream) {
    if (!in || !out || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < rows) return AC_E_INVALID_ARG;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    if (grid.y > 65535) return AC_E_INVALID_ARG;
    transpose_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(in), rows, cols, ld_in, reinterpret_cast<__half*>(out), ld_out);
    return acb::launched();
}

int ac_sd_conv_s2_dgrad_operand_f16(const float* dy, int B, int Ho, int Wo, int N, int H, int W, void* out, void* stream) {
    if (!dy || !out || B <= 0 || Ho <= 0 || Wo <= 0 || N <= 0 || H <= 0 || W <= 0) return AC_E_INVALID_ARG;
    const int Np = (N + 7) / 8 * 8;
    const long long total = (long long)B * H * W * 9 * (Np / 8);
    conv_s2_dgrad_operand_kernel<<<grid_for(total, 256, 16), 256, 0, (cudaStream_t)stream>>>(dy, B, Ho, Wo, N, Np, H, W, reinterpret_cast<__half*>(out));
    return acb::launched();
}

}  // extern "C"
