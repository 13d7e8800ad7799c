// nsr_device.cuh -- device building blocks of the Instant-NSR hot path (sm_100a).
//
// This translation unit is compiled with --fmad=false: every fused multiply-add below is an
// explicit fmaf(), so the elementwise arithmetic of the render core rounds exactly like the
// reference's un-fused eager torch ops, while the contractions the reference's own CUDA
// kernel gets from nvcc (cell position, corner blend) and the MLP dot products use FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acb {

constexpr int kLevels = 16;
constexpr int kHidden = 64;
constexpr int kSdfInPad = 36;   // 3 raw xyz + 32 hash features + 1 zero pad
constexpr int kColInPad = 24;   // 3 xyz + 3 normal + 15 geometry features + 3 zero pad

// Packed MLP blob (floats).  Every block starts on a 16 B boundary so rows can be read as float4.
constexpr int OFF_W0 = 0;                    // [64][36]  sdf layer 0, row j = hidden unit j
constexpr int OFF_B0 = OFF_W0 + 64 * 36;     // [64]
constexpr int OFF_W1T = OFF_B0 + 64;         // [64][16]  sdf layer 1, TRANSPOSED: row j = hidden unit j
constexpr int OFF_B1 = OFF_W1T + 64 * 16;    // [16]
constexpr int OFF_C0 = OFF_B1 + 16;          // [64][24]  colour layer 0
constexpr int OFF_C1 = OFF_C0 + 64 * 24;     // [64][64]  colour layer 1
constexpr int OFF_C2T = OFF_C1 + 64 * 64;    // [64][4]   colour layer 2, TRANSPOSED (col 3 = 0)
// Epilogue block, laid out exactly as the render kernel's __constant__ bank (copied with one D2D memcpy):
constexpr int OFF_EPI = OFF_C2T + 64 * 4;    // XB[64][4] = (wx, wy, wz, b0) | W1T[64][16] | B1[16] | C2T[64][4]
constexpr int EPI_XB = 0, EPI_W1T = EPI_XB + 64 * 4, EPI_B1 = EPI_W1T + 64 * 16, EPI_C2T = EPI_B1 + 16;
constexpr int EPI_FLOATS = EPI_C2T + 64 * 4;  // 1552
constexpr int BLOB_FLOATS = OFF_EPI + EPI_FLOATS;
static_assert(BLOB_FLOATS == 10848, "blob size is part of the C ABI");

// Experiment hook (scripts/build_variants.py): AC_TEX_MIN_SCALE routes the gathers of hashed levels with scale above the
// threshold through the texture pipe (tex1Dfetch on a linear float2 texture of the table: the same bits, another path
// through l1tex).  Measured (profiles/r02_gather_experiments.md): 1-2 % slower than ld.global.nc at every threshold.
#if defined(AC_TEX_MIN_SCALE)
__constant__ cudaTextureObject_t c_table_tex;
#endif

struct LevelMeta {
    uint32_t offset;   // first table entry of the level
    uint32_t size;     // entries in the level ("hashmap_size")
    uint32_t res1;     // resolution + 1 (dense stride)
    float scale;       // exp2f(level*S)*H - 1
    uint32_t hashed;   // 0: dense walk, 1: hashed, size is a power of two, 2: hashed, generic modulo
};

// Level geometry exactly as the reference kernel derives it per thread
// (encoder/hashencoder/src/hashencoder.cu:120-122, :54-70).
__device__ __forceinline__ LevelMeta make_level_meta(const int32_t* __restrict__ offsets, uint32_t level,
                                                     float S, uint32_t H, uint32_t D) {
    LevelMeta m;
    m.offset = (uint32_t)offsets[level];
    m.size = (uint32_t)(offsets[level + 1] - offsets[level]);
    m.scale = fmaf(exp2f((float)level * S), (float)H, -1.0f);   // nvcc contracts e*H-1 in the reference
    const uint32_t res = (uint32_t)ceilf(m.scale) + 1u;
    m.res1 = res + 1u;
    uint32_t stride = 1;
    for (uint32_t d = 0; d < D && stride <= m.size; ++d) stride *= m.res1;
    m.hashed = stride > m.size ? (((m.size & (m.size - 1u)) == 0u) ? 1u : 2u) : 0u;
    return m;
}

// Hashed levels whose size is not a power of two (never the case for the reference's configuration) take a real
// function call, so the 20-instruction integer modulo is not if-converted into the common path.
static __device__ __noinline__ uint32_t slot_modulo(uint32_t h, uint32_t size) { return h % size; }
__device__ __forceinline__ uint32_t wrap_slot(uint32_t h, const LevelMeta& m) {
    if (m.hashed == 1u) return h & (m.size - 1u);
    return slot_modulo(h, m.size);
}

// Level bodies with the level kind known at compile time (used when the offsets table has the reference layout:
// leading dense levels, then power-of-two hashed levels).  Bit-identical to grid_level_3d.
#ifndef AC_PAIR_LOADS
#define AC_PAIR_LOADS 0     // measured (profiles/r02_gather_experiments.md): bit-identical, but 14.8 ms/frame vs 14.1 ms -- off
#endif
template <bool HASHED>
__device__ __forceinline__ float2 grid_level_3d_k(const float2* __restrict__ table, const LevelMeta m, float x, float y, float z) {
    float px = fmaf(x, m.scale, 0.5f), py = fmaf(y, m.scale, 0.5f), pz = fmaf(z, m.scale, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    px -= fx; py -= fy; pz -= fz;
    const float qx = 1.0f - px, qy = 1.0f - py, qz = 1.0f - pz;
    const float2* __restrict__ t = table + m.offset;
    uint32_t s[8];
    if (!HASHED) {
        const uint32_t s1 = m.res1, s2 = m.res1 * m.res1;
        const uint32_t b = ix + iy * s1 + iz * s2;
        s[0] = b;          s[1] = b + 1u;          s[2] = b + s1;          s[3] = b + s1 + 1u;
        s[4] = b + s2;     s[5] = b + s2 + 1u;     s[6] = b + s2 + s1;     s[7] = b + s2 + s1 + 1u;
    } else {
        const uint32_t mask = m.size - 1u;
        const uint32_t hx0 = ix, hx1 = ix + 1u;
        const uint32_t hy0 = iy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = iz * 805459861u, hz1 = hz0 + 805459861u;
        s[0] = (hx0 ^ hy0 ^ hz0) & mask; s[1] = (hx1 ^ hy0 ^ hz0) & mask;
        s[2] = (hx0 ^ hy1 ^ hz0) & mask; s[3] = (hx1 ^ hy1 ^ hz0) & mask;
        s[4] = (hx0 ^ hy0 ^ hz1) & mask; s[5] = (hx1 ^ hy0 ^ hz1) & mask;
        s[6] = (hx0 ^ hy1 ^ hz1) & mask; s[7] = (hx1 ^ hy1 ^ hz1) & mask;
    }
    float2 v[8];
#if defined(AC_PROBE_PAIR)
    // TIMING PROBE ONLY (wrong values): every x-pair fetched with one 16-byte load -- the upper bound of what
    // merging even-x corner pairs can save.  Never part of a shipped build.
    if (HASHED && m.scale > AC_PROBE_PAIR) {
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(t + (s[k] | 1u)));
            v[k] = make_float2(q.x, q.y); v[k + 1] = make_float2(q.z, q.w);
        }
    } else
#endif
#if defined(AC_PROBE_SKIP)
    if (HASHED && m.scale > AC_PROBE_SKIP) {      // TIMING PROBE ONLY: fine levels cost nothing
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = make_float2(px, py);
    } else
#endif
#if defined(AC_TEX_MIN_SCALE)
    if (HASHED && m.scale > AC_TEX_MIN_SCALE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = tex1Dfetch<float2>(c_table_tex, (int)(m.offset + s[k]));
    } else
#endif
    {
#if AC_PAIR_LOADS
        // The two x-corners of a cell are neighbours in memory more often than not: slot, slot + 1 on a dense level, and
        // s, s ^ 1 on a hashed level whenever ix is even (the hash leaves x un-multiplied).  One 16-byte load of the aligned
        // entry pair that holds the first corner also delivers the second one when both sit in the same pair -- which, for
        // the reference's offsets table (hashed levels start at odd entries), happens for every even ix iff the table starts
        // 8 bytes past a 16-byte boundary (the host keeps such a copy).  Otherwise a second 8-byte load fetches it.  Same
        // table entries either way: results are bit-identical; the L1 data pipe sees up to 25 % fewer wavefronts.  MEASURED
        // SLOWER than eight 8-byte loads (the selects and the predicated second load cost more issue slots than the saved
        // wavefronts return: 14.8 vs 14.1 ms per frame with a shifted table, 15.4 ms without), so it is compiled out.
        const uint32_t b8 = (uint32_t)(reinterpret_cast<uintptr_t>(table) >> 3) & 1u;
        const float4* __restrict__ t16 = reinterpret_cast<const float4*>(table - b8);
        const uint32_t e_base = m.offset + b8;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const uint32_t e0 = e_base + s[k], e1 = e_base + s[k + 1];
            const float4 q = __ldg(t16 + (e0 >> 1));
            const float2 lo = make_float2(q.x, q.y), hi = make_float2(q.z, q.w);
            v[k] = (e0 & 1u) ? hi : lo;
            if ((e1 >> 1) == (e0 >> 1)) v[k + 1] = (e1 & 1u) ? hi : lo;
            else v[k + 1] = __ldg(t + s[k + 1]);
        }
#else
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(t + s[k]);
#endif
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = (((k & 1) ? px : qx) * ((k & 2) ? py : qy)) * ((k & 4) ? pz : qz);
        r.x = fmaf(w, v[k].x, r.x);
        r.y = fmaf(w, v[k].y, r.y);
    }
    return r;
}

// One level of the D=3, C=2 grid for a point already mapped to [0,1]^3
// (hashencoder.cu:124-166).  Returns the two blended channels.
__device__ __forceinline__ float2 grid_level_3d(const float2* __restrict__ table, const LevelMeta& m,
                                                float x, float y, float z) {
    float px = fmaf(x, m.scale, 0.5f), py = fmaf(y, m.scale, 0.5f), pz = fmaf(z, m.scale, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    px -= fx; py -= fy; pz -= fz;
    const float qx = 1.0f - px, qy = 1.0f - py, qz = 1.0f - pz;
    const float2* __restrict__ t = table + m.offset;
    uint32_t s[8];
    if (m.hashed == 0u) {
        const uint32_t s1 = m.res1, s2 = m.res1 * m.res1;
        const uint32_t b = ix + iy * s1 + iz * s2;     // < size for inputs in [0,1]: "% size" is a no-op
        s[0] = b;          s[1] = b + 1u;          s[2] = b + s1;          s[3] = b + s1 + 1u;
        s[4] = b + s2;     s[5] = b + s2 + 1u;     s[6] = b + s2 + s1;     s[7] = b + s2 + s1 + 1u;
    } else {
        const uint32_t hx0 = ix, hx1 = ix + 1u;
        const uint32_t hy0 = iy * 2654435761u, hy1 = (iy + 1u) * 2654435761u;
        const uint32_t hz0 = iz * 805459861u, hz1 = (iz + 1u) * 805459861u;
        s[0] = wrap_slot(hx0 ^ hy0 ^ hz0, m); s[1] = wrap_slot(hx1 ^ hy0 ^ hz0, m);
        s[2] = wrap_slot(hx0 ^ hy1 ^ hz0, m); s[3] = wrap_slot(hx1 ^ hy1 ^ hz0, m);
        s[4] = wrap_slot(hx0 ^ hy0 ^ hz1, m); s[5] = wrap_slot(hx1 ^ hy0 ^ hz1, m);
        s[6] = wrap_slot(hx0 ^ hy1 ^ hz1, m); s[7] = wrap_slot(hx1 ^ hy1 ^ hz1, m);
    }
    float2 v[8];
#if defined(AC_FINE_NO_ALLOCATE) && AC_FINE_NO_ALLOCATE
    if (m.hashed != 0u && m.scale > 200.0f) {      // fine hashed levels: no reuse inside an SM, keep them out of L1
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v[k].x), "=f"(v[k].y) : "l"(t + s[k]));
    } else
#endif
    {
#if AC_PAIR_LOADS
        // The two x-corners of a cell are neighbours in memory more often than not: slot, slot + 1 on a dense level, and
        // s, s ^ 1 on a hashed level whenever ix is even (the hash leaves x un-multiplied).  One 16-byte load of the aligned
        // entry pair that holds the first corner also delivers the second one when both sit in the same pair -- which, for
        // the reference's offsets table (hashed levels start at odd entries), happens for every even ix iff the table starts
        // 8 bytes past a 16-byte boundary (the host keeps such a copy).  Otherwise a second 8-byte load fetches it.  Same
        // table entries either way: results are bit-identical; the L1 data pipe sees up to 25 % fewer wavefronts.  MEASURED
        // SLOWER than eight 8-byte loads (the selects and the predicated second load cost more issue slots than the saved
        // wavefronts return: 14.8 vs 14.1 ms per frame with a shifted table, 15.4 ms without), so it is compiled out.
        const uint32_t b8 = (uint32_t)(reinterpret_cast<uintptr_t>(table) >> 3) & 1u;
        const float4* __restrict__ t16 = reinterpret_cast<const float4*>(table - b8);
        const uint32_t e_base = m.offset + b8;
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const uint32_t e0 = e_base + s[k], e1 = e_base + s[k + 1];
            const float4 q = __ldg(t16 + (e0 >> 1));
            const float2 lo = make_float2(q.x, q.y), hi = make_float2(q.z, q.w);
            v[k] = (e0 & 1u) ? hi : lo;
            if ((e1 >> 1) == (e0 >> 1)) v[k + 1] = (e1 & 1u) ? hi : lo;
            else v[k + 1] = __ldg(t + s[k + 1]);
        }
#else
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(t + s[k]);
#endif
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = (((k & 1) ? px : qx) * ((k & 2) ? py : qy)) * ((k & 4) ? pz : qz);
        r.x = fmaf(w, v[k].x, r.x);
        r.y = fmaf(w, v[k].y, r.y);
    }
    return r;
}

// Same arithmetic as grid_level_3d with ONE code path for dense and hashed levels (both slot formulas are
// a handful of integer ops; selecting between them is cheaper than carrying two bodies per level through
// the instruction cache of a 32-warp persistent kernel).  Results are bit-identical to grid_level_3d.
__device__ __forceinline__ float2 grid_level_3d_u(const float2* __restrict__ table, const LevelMeta m, float x, float y, float z) {
    float px = fmaf(x, m.scale, 0.5f), py = fmaf(y, m.scale, 0.5f), pz = fmaf(z, m.scale, 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
    px -= fx; py -= fy; pz -= fz;
    const float qx = 1.0f - px, qy = 1.0f - py, qz = 1.0f - pz;
    const bool hashed = m.hashed != 0u;
    const uint32_t my = hashed ? 2654435761u : m.res1, mz = hashed ? 805459861u : m.res1 * m.res1;
    const uint32_t y0 = iy * my, y1 = y0 + my, z0 = iz * mz, z1 = z0 + mz;
    const float2* __restrict__ t = table + m.offset;
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t cx = ix + (k & 1), cy = (k & 2) ? y1 : y0, cz = (k & 4) ? z1 : z0;
        uint32_t slot = hashed ? (cx ^ cy ^ cz) : (cx + cy + cz);
        if (m.hashed == 1u) slot &= m.size - 1u;
        else if (m.hashed == 2u) slot %= m.size;
        v[k] = __ldg(t + slot);
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float w = (((k & 1) ? px : qx) * ((k & 2) ? py : qy)) * ((k & 4) ? pz : qz);
        r.x = fmaf(w, v[k].x, r.x);
        r.y = fmaf(w, v[k].y, r.y);
    }
    return r;
}

// torch.nn.Softplus(beta=100, threshold=20): x when 100x > 20, else log1p(exp(100x))/100.
__device__ __forceinline__ float softplus100(float x) {
    const float t = x * 100.0f;
    return t > 20.0f ? x : log1pf(expf(t)) / 100.0f;
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// Hidden-layer variant: log1p(e^t)/100 = max(x,0) + (ln2/100) * log2(1 + 2^(-|t| log2 e)), two MUFU ops
// (ex2.approx, lg2.approx) instead of libdevice expf + log1pf (~45 instructions).  What matters for the SDF is
// the ABSOLUTE error of h: <= 2^-22 * ln2 / 100 ~ 1.7e-9, i.e. ~3e-9 on the signed distance after the 64-wide
// second layer -- 30x below fp32 rounding of the distance itself.  For t > 20 it returns x exactly like torch.
__device__ __forceinline__ float softplus100_mufu(float x) {
    float e, l;
    const float a = fabsf(x) * -144.26950408889634f;             // -|100 x| * log2(e)
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return fmaf(l, 0.006931471805599453f, fmaxf(x, 0.0f));
}

// sigmoid(t) with two MUFU ops (ex2.approx, rcp.approx): relative error ~2e-7.  Used for softplus'(a) = sigmoid(100 a) in
// the training backward, where it multiplies an upstream gradient (the libdevice expf + IEEE divide it replaces cost ~20
// instructions per hidden unit).
__device__ __forceinline__ float sigmoid_mufu(float t) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// Hash-encode a point given in world units (HashEncoder.forward, hashgrid.py:126-142):
// x01 = (x + bound) / (2*bound); out-of-range -> all-zero features (hashencoder.cu:94-119).
// Fills in[0..35] = {x, y, z, 32 features, 0} -- the SDF network's input row
// (models/instant_nsr.py:630-633: the RAW xyz is concatenated).
__device__ __forceinline__ void encode_point(const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                             float bound, float x, float y, float z, float (&in)[kSdfInPad]) {
    const float two_b = 2.0f * bound;
    const float u = (x + bound) / two_b, v = (y + bound) / two_b, w = (z + bound) / two_b;
    const bool oob = (u < 0.f) | (u > 1.f) | (v < 0.f) | (v > 1.f) | (w < 0.f) | (w > 1.f);
    in[0] = x; in[1] = y; in[2] = z; in[35] = 0.f;
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
        float2 f = make_float2(0.f, 0.f);
        if (!oob) f = grid_level_3d(table, lv[l], u, v, w);
        in[3 + 2 * l] = f.x;
        in[4 + 2 * l] = f.y;
    }
}

// SDF MLP 35 -> 64 (softplus100) -> 16 (models/instant_nsr.py:627-642).  `sw` is the blob in
// shared memory; every lane reads the same weight address (broadcast, conflict-free).
// FULL=false evaluates only output 0 (the signed distance).
template <bool FULL>
__device__ __forceinline__ void sdf_mlp(const float* __restrict__ sw, const float (&in)[kSdfInPad],
                                        float (&out)[FULL ? 16 : 1]) {
#pragma unroll
    for (int o = 0; o < (FULL ? 16 : 1); ++o) out[o] = sw[OFF_B1 + o];
#pragma unroll 2
    for (int j = 0; j < kHidden; ++j) {
        const float4* __restrict__ wr = reinterpret_cast<const float4*>(sw + OFF_W0 + j * kSdfInPad);
        float a = sw[OFF_B0 + j];
#pragma unroll
        for (int q = 0; q < kSdfInPad / 4; ++q) {
            const float4 w4 = wr[q];
            a = fmaf(w4.x, in[4 * q + 0], a);
            a = fmaf(w4.y, in[4 * q + 1], a);
            a = fmaf(w4.z, in[4 * q + 2], a);
            a = fmaf(w4.w, in[4 * q + 3], a);
        }
        const float h = softplus100_mufu(a);
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

template <bool FULL>
__device__ __forceinline__ void sdf_point(const float2* __restrict__ table, const LevelMeta* __restrict__ lv,
                                          const float* __restrict__ sw, float bound, float x, float y, float z,
                                          float (&out)[FULL ? 16 : 1]) {
    float in[kSdfInPad];
    encode_point(table, lv, bound, x, y, z, in);
    sdf_mlp<FULL>(sw, in, out);
}

// Colour MLP 21 -> 64 (relu) -> 64 (relu) -> 3 (sigmoid) (models/instant_nsr.py:644-663).
// bias0: optional [64] added to layer 0's pre-activation (use_viewdirs: the SH-encoded direction's contribution).
__device__ __forceinline__ void color_mlp(const float* __restrict__ sw, const float (&in)[kColInPad], float (&rgb)[3],
                                          const float* __restrict__ bias0 = nullptr) {
    float h1[kHidden];
#pragma unroll
    for (int j = 0; j < kHidden; ++j) {
        const float4* __restrict__ wr = reinterpret_cast<const float4*>(sw + OFF_C0 + j * kColInPad);
        float a = bias0 ? bias0[j] : 0.f;
#pragma unroll
        for (int q = 0; q < kColInPad / 4; ++q) {
            const float4 w4 = wr[q];
            a = fmaf(w4.x, in[4 * q + 0], a);
            a = fmaf(w4.y, in[4 * q + 1], a);
            a = fmaf(w4.z, in[4 * q + 2], a);
            a = fmaf(w4.w, in[4 * q + 3], a);
        }
        h1[j] = fmaxf(a, 0.f);
    }
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
    for (int j = 0; j < kHidden; ++j) {
        const float4* __restrict__ wr = reinterpret_cast<const float4*>(sw + OFF_C1 + j * kHidden);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < kHidden / 4; ++q) {
            const float4 w4 = wr[q];
            a = fmaf(w4.x, h1[4 * q + 0], a);
            a = fmaf(w4.y, h1[4 * q + 1], a);
            a = fmaf(w4.z, h1[4 * q + 2], a);
            a = fmaf(w4.w, h1[4 * q + 3], a);
        }
        a = fmaxf(a, 0.f);
        const float4 c = *reinterpret_cast<const float4*>(sw + OFF_C2T + j * 4);
        o0 = fmaf(c.x, a, o0);
        o1 = fmaf(c.y, a, o1);
        o2 = fmaf(c.z, a, o2);
    }
    rgb[0] = sigmoidf(o0); rgb[1] = sigmoidf(o1); rgb[2] = sigmoidf(o2);
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ---- warp-level scans over <=128 values staged in shared memory (4 consecutive per lane) ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float& total) {
    float s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s = t * s;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    const float e = __shfl_up_sync(0xffffffffu, s, 1);
    return lane == 0 ? 1.0f : e;
}
__device__ __forceinline__ float warp_excl_sum(float v, int lane) {
    float s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s = t + s;
    }
    const float e = __shfl_up_sync(0xffffffffu, s, 1);
    return lane == 0 ? 0.0f : e;
}

// torch.linspace(0, 1, n)[i] on float32 (symmetric two-sided formula; the upper half is
// evaluated with a fused multiply-add, which is what torch's CPU and CUDA kernels produce).
__device__ __forceinline__ float linspace01(int i, int n) {
    const float step = 1.0f / (float)(n - 1);
    return i < n / 2 ? step * (float)i : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

struct Ray {
    float ox, oy, oz, dx, dy, dz;
};
__device__ __forceinline__ void ray_point(const Ray& r, float t, float& x, float& y, float& z) {
    x = r.ox + r.dx * t; y = r.oy + r.dy * t; z = r.oz + r.dz * t;      // mul, then add (no FMA: --fmad=false)
}

// near/far against the cube [-bound,bound]^3 (models/instant_nsr.py:58-77, type='cube').
__device__ __forceinline__ void ray_box(const Ray& r, float bound, float& near, float& far) {
    const float o[3] = {r.ox, r.oy, r.oz}, d[3] = {r.dx, r.dy, r.dz};
    near = -INFINITY; far = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float den = d[k] + 1e-15f;
        const float t0 = (-bound - o[k]) / den, t1 = (bound - o[k]) / den;
        const float lo = t0 < t1 ? t0 : t1, hi = t0 > t1 ? t0 : t1;
        near = lo > near ? lo : near;       // torch.max / torch.min over the 3 axes
        far = hi < far ? hi : far;
    }
    near = fmaxf(near, 0.05f);
}

// One importance round on a warp-owned ray (up_sample :410-459, sample_pdf :21-55 det=True).
// zs/sdfs hold T sorted depths and their SDF; ta/tb are T-float scratch rows.
// Step 1: per-interval alpha into ta[0..T-2].
__device__ __forceinline__ void interval_alpha(const Ray& r, const float* zs, const float* sdfs, float* ta, int T,
                                               float inv_s, int lane) {
    const int nI = T - 1;
    for (int k = lane; k < nI; k += 32) {
        const float z0 = zs[k], z1 = zs[k + 1], s0 = sdfs[k], s1 = sdfs[k + 1];
        float x, y, z;
        ray_point(r, z0, x, y, z);
        const float r0 = sqrtf(x * x + y * y + z * z);
        ray_point(r, z1, x, y, z);
        const float r1 = sqrtf(x * x + y * y + z * z);
        const float inside = ((r0 < 1.0f) | (r1 < 1.0f)) ? 1.0f : 0.0f;
        const float slope = (s1 - s0) / (z1 - z0 + 1e-5f);
        const float prev = k == 0 ? 0.0f : (s0 - sdfs[k - 1]) / (z0 - zs[k - 1] + 1e-5f);
        float c = fminf(prev, slope);
        c = fminf(fmaxf(c, -1e3f), 0.0f) * inside;
        const float dist = z1 - z0, mid = (s0 + s1) * 0.5f;
        const float half = c * dist * 0.5f;
        const float c0 = sigmoidf((mid - half) * inv_s), c1 = sigmoidf((mid + half) * inv_s);
        ta[k] = (c0 - c1 + 1e-5f) / (c0 + 1e-5f);
    }
    __syncwarp();
}

// Step 2: alpha (ta) -> weights -> pdf -> cdf (tb) -> 16 deterministic inverse-CDF samples.
// On return lanes 0..15 hold the new depths (ascending) and the CDF bin (below, above) of each.
__device__ __forceinline__ void importance_from_alpha(const float* zs, float* ta, float* tb, int T, int lane,
                                                      float& z_new, int& below, int& above) {
    const int nI = T - 1;
    // weights = alpha * exclusive_cumprod(1 - alpha + 1e-7); pdf numerators w + 1e-5
    float a[4], f[4];
    const int base = lane * 4;
    float p = 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = base + i < nI ? ta[base + i] : 0.0f;
        f[i] = base + i < nI ? (1.0f - a[i] + 1e-7f) : 1.0f;
        p *= f[i];
    }
    float total;
    float run = warp_excl_prod(p, lane, total);
    float wsum = 0.0f, w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i] = base + i < nI ? (a[i] * run + 1e-5f) : 0.0f;
        wsum += w[i];
        run *= f[i];
    }
    const float denom = warp_sum(wsum);
    float local = 0.0f, pdf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pdf[i] = w[i] / denom;
        local += pdf[i];
    }
    float acc = warp_excl_sum(local, lane);
    __syncwarp();
    if (lane == 0) tb[0] = 0.0f;                 // cdf = cat([0, cumsum(pdf)])  -> T entries in tb
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc += pdf[i];
        if (base + i < nI) tb[base + i + 1] = acc;
    }
    __syncwarp();
    z_new = 0.0f; below = 0; above = 0;
    if (lane < 16) {
        const float u = (float)(2 * lane + 1) / 32.0f;          // linspace(1/32, 31/32, 16), exact
        int lo = 0, hi = T;                                      // searchsorted(cdf, u, right=True)
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (tb[m] <= u) lo = m + 1; else hi = m;
        }
        below = lo - 1 > 0 ? lo - 1 : 0;
        above = lo < T - 1 ? lo : T - 1;
        const float cb = tb[below], ca = tb[above], zb = zs[below], za = zs[above];
        float den = ca - cb;
        den = den < 1e-5f ? 1.0f : den;
        const float t = (u - cb) / den;
        z_new = zb + t * (za - zb);
    }
    __syncwarp();
}

__device__ __forceinline__ void importance_round(const Ray& r, const float* zs, const float* sdfs, float* ta,
                                                 float* tb, int T, float inv_s, int lane, float& z_new,
                                                 int& below, int& above) {
    interval_alpha(r, zs, sdfs, ta, T, inv_s, lane);
    importance_from_alpha(zs, ta, tb, T, lane, z_new, below, above);
}

// Merge 16 new ascending depths (lanes 0..15) into the T sorted ones (torch.sort of the
// concatenation, cat_z_vals :461-475).  Ties keep the old depth first.  pos_old[i] for the
// lane's old elements k = lane + 32*i, pos_new for the lane's new element.
__device__ __forceinline__ void merge_positions(const float* zs, int T, float z_new, int lane, int (&pos_old)[4],
                                                int& pos_new) {
    float zo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = lane + 32 * i;
        zo[i] = k < T ? zs[k] : INFINITY;
        pos_old[i] = k;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float zj = __shfl_sync(0xffffffffu, z_new, j);
#pragma unroll
        for (int i = 0; i < 4; ++i) pos_old[i] += zj < zo[i] ? 1 : 0;
    }
    int lo = 0, hi = T;                           // #old <= z_new  (upper bound)
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (zs[m] <= z_new) lo = m + 1; else hi = m;
    }
    pos_new = lane + lo;
}

}  // namespace acb
