// nsr_shade_tc.cu -- the differentiable half of NeRFRenderer.run for the training path (sm_100a): everything AFTER the
// 7-point SDF stencil has been evaluated (models/instant_nsr.py:210-299), forward and backward, as two kernels.
//
//   forward  (ac_nsr_shade_forward):  finite-difference normal -> colour MLP (tcgen05) -> NeuS alpha -> transmittance scan ->
//            image / weight_sum / normal map / depth / eikonal partials (+ per-sample weights, colour, alpha)
//   backward (ac_nsr_shade_backward): reverse scan of the compositing, alpha / softplus / sigmoid / normalisation algebra,
//            colour-MLP recompute + data gradients on tcgen05 (row = sample tiles, transposed weight tiles), emitting
//            g_centre [M,16] and g_fd [6,M] -- exactly what ac_nsr_sdf_backward_stencil consumes -- plus the fp16 per-sample
//            terms of the colour-MLP weight gradients, which ONE ac_sd_gemm_f16 launch (TMA + tcgen05, split-K) reduces
//            over the M samples:   [da2 ; da1 ; dz2] [136 x M]  x  [h1 ; cin ; h2]^T [M x 160].
//
// In the reference (and in round 1 here) this half is ~40 eager torch ops + autograd per 4096-ray patch.
//
// Layout: a warp owns a ray, lane = sample of a 32-sample block, four warps = one 128-row MMA tile (nsr_tc_group.cuh).
// Gradients that pass through the colour MLP are multiplied by a power-of-two `scale` (device scalar, chosen by the caller
// from max|g_rgb|) before they become fp16 tensor-core operands and divided out again in fp32.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"
#include "tc05.cuh"
#include "nsr_tc_group.cuh"

using namespace acb;

namespace {

constexpr int kGroupsH = 4;                     // 4-warp groups per CTA: 512 threads, <= 128 registers
constexpr int kWarpsH = 4 * kGroupsH;
constexpr int kMaxTH = 128;

// shared memory map (bytes).  Weight tiles, each (hi | lo) in the UMMA K-major no-swizzle layout:
constexpr uint32_t HB_C0_HI = 0, HB_C0_LO = 4096;            // colour layer 0: [n = 64 units][k = 32 inputs]
constexpr uint32_t HB_C1_HI = 8192, HB_C1_LO = 16384;        // colour layer 1: [64][64]
constexpr uint32_t HB_C1T_HI = 24576, HB_C1T_LO = 32768;     // its transpose (data gradient): [n = 64 inputs][k = 64 units]
constexpr uint32_t HB_C0T_HI = 40960, HB_C0T_LO = 45056;     // layer 0 transposed: [n = 32 inputs][k = 64 units], chunk stride 512
constexpr uint32_t HB_BYTES = 49152;
constexpr size_t SH_EPI = 0;                                  // C2T [64][4] floats
constexpr size_t SH_B = 1024;
constexpr size_t SH_A = SH_B + HB_BYTES;                      // per group 32 KB: K = 64 activations as (hi | lo) fp16 in the forward
constexpr size_t kGroupA = 32768;
constexpr size_t SH_BARS = SH_A + (size_t)kGroupsH * kGroupA;
constexpr size_t SH_TOTAL = SH_BARS + kGroupsH * 8 + 16;

struct ShadeParams {
    ac_nsr_shade_args a;
    const float* blob;
    const float* variance;
};

__device__ __forceinline__ void stage_b_tile_n(unsigned char* bhi, unsigned char* blo, int n, int k, float w, int chunk_stride) {
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const int at = (k >> 3) * chunk_stride + n * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(bhi + at) = h;
    *reinterpret_cast<__half*>(blo + at) = l;
}

// One-time staging shared by both kernels; returns after the CTA-wide barrier.
__device__ __forceinline__ uint32_t stage_common(unsigned char* smem, const float* __restrict__ blob, bool transposed) {
    float* c2t = reinterpret_cast<float*>(smem + SH_EPI);
    unsigned char* bt = smem + SH_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SH_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroupsH);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 4; i += blockDim.x) c2t[i] = __ldg(blob + OFF_C2T + i);
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int n = i >> 5, k = i & 31;
        const float w = k < kColInPad ? __ldg(blob + OFF_C0 + n * kColInPad + k) : 0.f;
        stage_b_tile_n(bt + HB_C0_HI, bt + HB_C0_LO, n, k, w, 1024);
        if (transposed) stage_b_tile_n(bt + HB_C0T_HI, bt + HB_C0T_LO, k, n, w, 512);      // [n = input k][k = unit n]
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i >> 6, k = i & 63;
        const float w = __ldg(blob + OFF_C1 + n * kHidden + k);
        stage_b_tile_n(bt + HB_C1_HI, bt + HB_C1_LO, n, k, w, 1024);
        if (transposed) stage_b_tile_n(bt + HB_C1T_HI, bt + HB_C1T_LO, k, n, w, 1024);
    }
    if (threadIdx.x == 0) {
        for (int g = 0; g < kGroupsH; ++g) tc05::mbar_init(bars + g, 1);
        tc05::fence_mbar_init();
    }
    if (warp == 0) tc05::tmem_alloc<64 * kGroupsH>(tmem_slot);
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    return *tmem_slot;
}

__device__ __forceinline__ Group make_group(unsigned char* smem, uint32_t tmem_base) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, group = warp >> 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SH_BARS);
    Group g;
    g.a = smem + SH_A + (size_t)group * kGroupA;
    g.a_s = tc05::smem_u32(g.a);
    g.b_s = tc05::smem_u32(smem + SH_B);
    g.bar = bars + group;
    g.phase = 0;
    g.row = (warp & 3) * 32 + lane;
    g.tmem = tmem_base + (uint32_t)group * 64u + ((uint32_t)((warp & 3) * 32) << 16);
    g.bar_id = 1 + group;
    g.std_layout = true;
    return g;
}

// ---- per-sample geometry: normal from the stencil, NeuS alpha (models/instant_nsr.py:214-250) ----
struct SampleGeo {
    float px, py, pz;      // section point (clamped)
    float gx, gy, gz, gn;  // finite-difference gradient and its norm
    float nx, ny, nz;      // normal = g / (1e-5 + |g|)
    float sdf, gap;
    float cosv, it, hs, c0, c1, ratio, alpha;
};

__device__ __forceinline__ void sample_geo(const ac_nsr_shade_args& a, size_t m, uint32_t M, float eps, const Ray& r, float gap, float inv_s,
                                           SampleGeo& s) {
    s.px = a.points[3 * m]; s.py = a.points[3 * m + 1]; s.pz = a.points[3 * m + 2];
    s.sdf = a.centre[16 * m];
    const float f0 = a.fd[m], f1 = a.fd[(size_t)M + m], f2 = a.fd[2 * (size_t)M + m], f3 = a.fd[3 * (size_t)M + m],
                f4 = a.fd[4 * (size_t)M + m], f5 = a.fd[5 * (size_t)M + m];
    s.gx = 0.5f * (f0 - f1) / eps; s.gy = 0.5f * (f2 - f3) / eps; s.gz = 0.5f * (f4 - f5) / eps;
    s.gn = sqrtf(s.gx * s.gx + s.gy * s.gy + s.gz * s.gz);
    const float inv = 1e-5f + s.gn;
    s.nx = s.gx / inv; s.ny = s.gy / inv; s.nz = s.gz / inv;
    s.gap = gap;
    s.cosv = r.dx * s.nx + r.dy * s.ny + r.dz * s.nz;
    const float car = a.cos_anneal_ratio;
    s.it = -(softplus100(-s.cosv * 0.5f + 0.5f) * (1.0f - car) + softplus100(-s.cosv) * car);
    s.hs = s.it * gap * 0.5f;
    s.c0 = sigmoidf((s.sdf - s.hs) * inv_s); s.c1 = sigmoidf((s.sdf + s.hs) * inv_s);
    s.ratio = (s.c0 - s.c1 + 1e-5f) / (s.c0 + 1e-5f);
    s.alpha = clampf(s.ratio, 0.0f, 1.0f);
}

// cin = (x, y, z, nx, ny, nz, 15 geometry features, 0 0 0) -> this thread's A row (hi | lo), K = 32
__device__ __forceinline__ void write_cin_row(Group& g, const float (&cin)[24]) {
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
}

// D[128 x N] = A[128 x 64] (fp16, 8 chunks) * B(hi, lo)[N x 64]^T
template <uint32_t N>
__device__ __forceinline__ void issue_k64_x2(uint32_t tmem_d, uint32_t a_s, uint32_t bhi_s, uint32_t blo_s, uint32_t b_chunk_stride) {
    constexpr uint32_t idesc = tc05::idesc_f16(128, N);
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        const uint64_t ad = tc05::smem_desc(a_s + s * 4096u, 2048u, 128u);
        const uint64_t bh = tc05::smem_desc(bhi_s + s * 2u * b_chunk_stride, b_chunk_stride, 128u);
        const uint64_t bl = tc05::smem_desc(blo_s + s * 2u * b_chunk_stride, b_chunk_stride, 128u);
        tc05::mma_f16(tmem_d, ad, bh, idesc, s);
        tc05::mma_f16(tmem_d, ad, bl, idesc, 1u);
    }
}

// D[128 x 64] = A(hi, lo)[128 x 64] * B(hi, lo)[64 x 64]^T with the three fp16 partial products (A lo at +16 KB)
__device__ __forceinline__ void issue_k64_x3(uint32_t tmem_d, uint32_t a_s, uint32_t bhi_s, uint32_t blo_s) {
    constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
#pragma unroll
    for (uint32_t s = 0; s < 4; ++s) {
        const uint64_t ah = tc05::smem_desc(a_s + s * 4096u, 2048u, 128u);
        const uint64_t al = tc05::smem_desc(a_s + 16384u + s * 4096u, 2048u, 128u);
        const uint64_t bh = tc05::smem_desc(bhi_s + s * 2048u, 1024u, 128u);
        const uint64_t bl = tc05::smem_desc(blo_s + s * 2048u, 1024u, 128u);
        tc05::mma_f16(tmem_d, ah, bh, idesc, s);
        tc05::mma_f16(tmem_d, al, bh, idesc, 1u);
        tc05::mma_f16(tmem_d, ah, bl, idesc, 1u);
    }
}

__device__ __forceinline__ void load_ray(const ac_nsr_shade_args& a, uint32_t ray, Ray& r, float& near, float& span) {
    r.ox = a.rays_o[3 * ray]; r.oy = a.rays_o[3 * ray + 1]; r.oz = a.rays_o[3 * ray + 2];
    r.dx = a.rays_d[3 * ray]; r.dy = a.rays_d[3 * ray + 1]; r.dz = a.rays_d[3 * ray + 2];
    float far;
    ray_box(r, a.bound, near, far);
    span = far - near;
}

// ============================================================ forward ============================================================
__global__ void __launch_bounds__(kWarpsH * 32, 1) shade_forward_kernel(const ShadeParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tmem_base = stage_common(smem, p.blob, false);
    Group g = make_group(smem, tmem_base);
    const float* c2t = reinterpret_cast<const float*>(smem + SH_EPI);
    const ac_nsr_shade_args& a = p.a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, group = warp >> 2, wq = warp & 3;
    const uint32_t T = a.n_samples, M = a.n_rays * T;
    const float inv_s = clampf(expf(p.variance[0] * 10.0f), 1e-6f, 1e6f);
    const float eps = a.eps;
    const uint32_t n_quads = (a.n_rays + 3) / 4;

    for (uint32_t quad = blockIdx.x + gridDim.x * group; quad < n_quads; quad += gridDim.x * kGroupsH) {
        const uint32_t ray_raw = quad * 4 + wq;
        const bool ray_ok = ray_raw < a.n_rays;
        const uint32_t ray = ray_ok ? ray_raw : a.n_rays - 1;
        Ray r; float near, span;
        load_ray(a, ray, r, near, span);
        const float sample_dist = span / (float)a.num_steps;
        const float* zr = a.z_vals + (size_t)ray * T;
        float carry = 1.0f;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_nx = 0.f, acc_ny = 0.f, acc_nz = 0.f, acc_w = 0.f, acc_d = 0.f, en = 0.f, ed = 0.f;
        for (uint32_t k0 = 0; k0 < T; k0 += 32) {
            const bool live = k0 + lane < T;
            const uint32_t k = min(k0 + lane, T - 1);
            const size_t m = (size_t)ray * T + k;
            const float zk = zr[k];
            const float gap = k < T - 1 ? zr[k + 1] - zk : sample_dist;
            SampleGeo s;
            sample_geo(a, m, M, eps, r, gap, inv_s, s);
            float cin[24];
            cin[0] = s.px; cin[1] = s.py; cin[2] = s.pz; cin[3] = s.nx; cin[4] = s.ny; cin[5] = s.nz;
            {
                const float4* c4 = reinterpret_cast<const float4*>(a.centre + 16 * m);
                const float4 q0 = c4[0], q1 = c4[1], q2 = c4[2], q3 = c4[3];
                cin[6] = q0.y; cin[7] = q0.z; cin[8] = q0.w; cin[9] = q1.x; cin[10] = q1.y; cin[11] = q1.z; cin[12] = q1.w;
                cin[13] = q2.x; cin[14] = q2.y; cin[15] = q2.z; cin[16] = q2.w; cin[17] = q3.x; cin[18] = q3.y; cin[19] = q3.z; cin[20] = q3.w;
            }
            cin[21] = cin[22] = cin[23] = 0.f;
            write_cin_row(g, cin);
            group_mma_round(g, [&] { issue_k32_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C0_HI, g.b_s + HB_C0_LO); });
#pragma unroll 1
            for (int qtr = 0; qtr < 4; ++qtr) {
                float acc[16];
                tc05::tmem_ld16(g.tmem + qtr * 16, acc);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 hi, lo;
                    tc05::split_f16x2(fmaxf(acc[8 * half + 0], 0.f), fmaxf(acc[8 * half + 1], 0.f), hi.x, lo.x);
                    tc05::split_f16x2(fmaxf(acc[8 * half + 2], 0.f), fmaxf(acc[8 * half + 3], 0.f), hi.y, lo.y);
                    tc05::split_f16x2(fmaxf(acc[8 * half + 4], 0.f), fmaxf(acc[8 * half + 5], 0.f), hi.z, lo.z);
                    tc05::split_f16x2(fmaxf(acc[8 * half + 6], 0.f), fmaxf(acc[8 * half + 7], 0.f), hi.w, lo.w);
                    *reinterpret_cast<uint4*>(g.a + (2 * qtr + half) * 2048 + g.row * 16) = hi;
                    *reinterpret_cast<uint4*>(g.a + 16384 + (2 * qtr + half) * 2048 + g.row * 16) = lo;
                }
            }
            group_mma_round(g, [&] { issue_k64_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C1_HI, g.b_s + HB_C1_LO); });
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 1
            for (int qtr = 0; qtr < 4; ++qtr) {
                float acc[16];
                tc05::tmem_ld16(g.tmem + qtr * 16, acc);
                const float4* __restrict__ c2 = reinterpret_cast<const float4*>(c2t + qtr * 16 * 4);
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const float h2 = fmaxf(acc[jj], 0.f);
                    const float4 w = c2[jj];
                    o0 = fmaf(w.x, h2, o0); o1 = fmaf(w.y, h2, o1); o2 = fmaf(w.z, h2, o2);
                }
            }
            const float cr = sigmoidf(o0), cg = sigmoidf(o1), cb = sigmoidf(o2);
            const float alpha = live ? s.alpha : 0.0f;
            const float pn = sqrtf(s.px * s.px + s.py * s.py + s.pz * s.pz);
            if (live && pn < 1.2f) { en += (s.gn - 1.0f) * (s.gn - 1.0f); ed += 1.0f; }
            float blk;
            const float tr = warp_excl_prod(live ? (1.0f - alpha + 1e-7f) : 1.0f, lane, blk) * carry;
            carry *= blk;
            const float w = alpha * tr;
            if (live) {
                acc_r += w * cr; acc_g += w * cg; acc_b += w * cb;
                acc_nx += w * s.nx; acc_ny += w * s.ny; acc_nz += w * s.nz;
                acc_w += w;
                acc_d += w * clampf((zk - near) / span, 0.0f, 1.0f);
                if (ray_ok) {
                    a.weights[m] = w; a.pts_alpha[m] = alpha;
                    a.pts_color[3 * m] = cr; a.pts_color[3 * m + 1] = cg; a.pts_color[3 * m + 2] = cb;
                }
            }
        }
        acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
        acc_nx = warp_sum(acc_nx); acc_ny = warp_sum(acc_ny); acc_nz = warp_sum(acc_nz);
        acc_w = warp_sum(acc_w); acc_d = warp_sum(acc_d); en = warp_sum(en); ed = warp_sum(ed);
        if (lane == 0 && ray_ok) {
            float bg[3] = {1.f, 1.f, 1.f};
            if (a.bg_color) { bg[0] = a.bg_color[3 * ray]; bg[1] = a.bg_color[3 * ray + 1]; bg[2] = a.bg_color[3 * ray + 2]; }
            const float rest = 1.0f - acc_w;
            a.rgb[3 * ray] = acc_r + rest * bg[0]; a.rgb[3 * ray + 1] = acc_g + rest * bg[1]; a.rgb[3 * ray + 2] = acc_b + rest * bg[2];
            a.depth[ray] = acc_d; a.weight_sum[ray] = acc_w;
            a.normal[3 * ray] = acc_nx; a.normal[3 * ray + 1] = acc_ny; a.normal[3 * ray + 2] = acc_nz;
            a.eik_partial[2 * ray] = en; a.eik_partial[2 * ray + 1] = ed;
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64 * kGroupsH>(tmem_base);
}

// eikonal = sum(num) / (sum(den) + 1e-5) over all rays (models/instant_nsr.py:266-272); out[0] = eikonal, out[1] = sum(den)
__global__ void __launch_bounds__(1024) shade_eik_reduce_kernel(const float* __restrict__ partial, uint32_t n, float* __restrict__ out) {
    __shared__ float sn[32], sd[32];
    float num = 0.f, den = 0.f;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { num += partial[2 * i]; den += partial[2 * i + 1]; }
    num = warp_sum(num); den = warp_sum(den);
    if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x < 32) {
        num = sn[threadIdx.x]; den = sd[threadIdx.x];
        num = warp_sum(num); den = warp_sum(den);
        if (threadIdx.x == 0) { out[0] = num / (den + 1e-5f); out[1] = den; }
    }
}

// scale = 2^floor(log2(target / max|x|)) (device scalar; 1 when x is all zero) -- the fp16 operand scale of the backward
__global__ void __launch_bounds__(1024) absmax_scale_kernel(const float* __restrict__ x, uint32_t n, float target, float* __restrict__ out) {
    __shared__ float sm[32];
    float mx = 0.f;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = sm[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) {
            float e = 0.f;
            if (mx > 0.f && isfinite(mx)) e = fminf(fmaxf(floorf(log2f(target / mx)), -100.f), 100.f);
            out[0] = exp2f(e);
        }
    }
}

// ============================================================ backward ============================================================
__global__ void __launch_bounds__(kWarpsH * 32, 1) shade_backward_kernel(const ShadeParams p, const ac_nsr_shade_grads gr) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t tmem_base = stage_common(smem, p.blob, true);
    Group g = make_group(smem, tmem_base);
    const float* c2t = reinterpret_cast<const float*>(smem + SH_EPI);
    const ac_nsr_shade_args& a = p.a;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, group = warp >> 2, wq = warp & 3;
    const uint32_t T = a.n_samples, M = a.n_rays * T;
    const float ex = expf(p.variance[0] * 10.0f);
    const float inv_s = clampf(ex, 1e-6f, 1e6f);
    const bool s_live = ex >= 1e-6f && ex <= 1e6f;            // clip passes the gradient inside [min, max]
    const float eps = a.eps;
    const uint32_t n_quads = (a.n_rays + 3) / 4;
    const float scale = gr.scale[0], inv_scale = 1.0f / scale;
    const float g_eik = gr.g_eikonal ? gr.g_eikonal[0] : 0.f;
    const float eik_den = gr.eik_out[1] + 1e-5f;
    const int nb = (int)((T + 31) / 32);
    __half* tA = reinterpret_cast<__half*>(gr.terms_a);
    __half* tB = reinterpret_cast<__half*>(gr.terms_b);
    const size_t ld = (size_t)gr.terms_ld;
    float ds_acc = 0.f;                                       // d loss / d inv_s, summed over this thread's samples
    float b1_acc[16];                                         // column sums of g_centre (+ all of g_fd in column 0) = d loss / d b1
#pragma unroll
    for (int q = 0; q < 16; ++q) b1_acc[q] = 0.f;

    for (uint32_t quad = blockIdx.x + gridDim.x * group; quad < n_quads; quad += gridDim.x * kGroupsH) {
        const uint32_t ray_raw = quad * 4 + wq;
        const bool ray_ok = ray_raw < a.n_rays;
        const uint32_t ray = ray_ok ? ray_raw : a.n_rays - 1;
        Ray r; float near, span;
        load_ray(a, ray, r, near, span);
        const float sample_dist = span / (float)a.num_steps;
        const float* zr = a.z_vals + (size_t)ray * T;
        float gi[3] = {gr.g_rgb[3 * ray], gr.g_rgb[3 * ray + 1], gr.g_rgb[3 * ray + 2]};
        float bg[3] = {1.f, 1.f, 1.f};
        if (a.bg_color) { bg[0] = a.bg_color[3 * ray]; bg[1] = a.bg_color[3 * ray + 1]; bg[2] = a.bg_color[3 * ray + 2]; }
        float g_ws = gr.g_weight_sum ? gr.g_weight_sum[ray] : 0.f;
        if (gr.wsum_gt) {
            // fused opacity term of the trainer (stylize.py:187-193): opacity_weight * smooth_l1(clamp(ws), clamp(ws_gt)), mean over
            // the launch's rays; d/d ws = weight / n * (|x| < 1 ? x : sign x) inside the clamp
            const float ws = a.weight_sum[ray], wg = gr.wsum_gt[ray];
            const float x = clampf(ws, 0.f, 1.f) - clampf(wg, 0.f, 1.f);
            const float d = fabsf(x) < 1.0f ? x : (x > 0.f ? 1.0f : -1.0f);
            if (ws >= 0.f && ws <= 1.f) g_ws += gr.opacity_weight / (float)a.n_rays * d;
            if (gr.opacity_loss && lane == 0 && ray_ok) {
                const float l = fabsf(x) < 1.0f ? 0.5f * x * x : fabsf(x) - 0.5f;
                if (l != 0.f) atomicAdd(gr.opacity_loss, gr.opacity_weight / (float)a.n_rays * l);
            }
        }
        float gnm[3] = {0.f, 0.f, 0.f};
        if (gr.g_normal) { gnm[0] = gr.g_normal[3 * ray]; gnm[1] = gr.g_normal[3 * ray + 1]; gnm[2] = gr.g_normal[3 * ray + 2]; }
        const float g_dep = gr.g_depth ? gr.g_depth[ray] : 0.f;
        if (!ray_ok) { gi[0] = gi[1] = gi[2] = 0.f; g_ws = 0.f; gnm[0] = gnm[1] = gnm[2] = 0.f; }

        // ---- phase 1: d loss / d alpha for every sample of the ray (reverse scan of the compositing) ----
        float dalpha[4], trans[4];
        {
            float prod[4], gw[4], al[4], tot[4];
            float carry = 1.0f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                prod[b] = 0.f; gw[b] = 0.f; al[b] = 0.f; tot[b] = 0.f; trans[b] = 1.f; dalpha[b] = 0.f;
                if (b < nb) {
                    const uint32_t k = 32u * b + lane;
                    const bool live = k < T;
                    const size_t m = (size_t)ray * T + min(k, T - 1);
                    const float w = live ? a.weights[m] : 0.f;
                    al[b] = live ? a.pts_alpha[m] : 0.f;
                    float v = 0.f;
                    if (live) {
                        v = gi[0] * (a.pts_color[3 * m] - bg[0]) + gi[1] * (a.pts_color[3 * m + 1] - bg[1]) + gi[2] * (a.pts_color[3 * m + 2] - bg[2]) + g_ws;
                        if (gr.g_normal || gr.g_depth) {
                            SampleGeo s;
                            const float zk = zr[k];
                            sample_geo(a, m, M, eps, r, k < T - 1 ? zr[k + 1] - zk : sample_dist, inv_s, s);
                            v += gnm[0] * s.nx + gnm[1] * s.ny + gnm[2] * s.nz + g_dep * clampf((zk - near) / span, 0.0f, 1.0f);
                        }
                    }
                    gw[b] = v;
                    prod[b] = v * w;
                    float blk;
                    trans[b] = warp_excl_prod(live ? (1.0f - al[b] + 1e-7f) : 1.0f, lane, blk) * carry;
                    carry *= blk;
                    tot[b] = warp_sum(prod[b]);
                }
            }
            float after = 0.f;
#pragma unroll
            for (int b = 3; b >= 0; --b) {
                if (b < nb) {
                    const float incl = warp_excl_sum(prod[b], lane) + prod[b];
                    const float suffix = (tot[b] - incl) + after;          // sum over samples behind this one
                    dalpha[b] = gw[b] * trans[b] - suffix / (1.0f - al[b] + 1e-7f);
                    after += tot[b];
                }
            }
        }

        // ---- phase 2: per block, colour MLP recompute + backward, alpha / normal algebra, outputs ----
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const uint32_t k = min(32u * b + lane, T - 1);
            const bool live = 32u * b + lane < T && ray_ok;
            const size_t m = (size_t)ray * T + k;
            const float zk = zr[k];
            SampleGeo s;
            sample_geo(a, m, M, eps, r, k < T - 1 ? zr[k + 1] - zk : sample_dist, inv_s, s);
            float cin[24];
            cin[0] = s.px; cin[1] = s.py; cin[2] = s.pz; cin[3] = s.nx; cin[4] = s.ny; cin[5] = s.nz;
            {
                const float4* c4 = reinterpret_cast<const float4*>(a.centre + 16 * m);
                const float4 q0 = c4[0], q1 = c4[1], q2 = c4[2], q3 = c4[3];
                cin[6] = q0.y; cin[7] = q0.z; cin[8] = q0.w; cin[9] = q1.x; cin[10] = q1.y; cin[11] = q1.z; cin[12] = q1.w;
                cin[13] = q2.x; cin[14] = q2.y; cin[15] = q2.z; cin[16] = q2.w; cin[17] = q3.x; cin[18] = q3.y; cin[19] = q3.z; cin[20] = q3.w;
            }
            cin[21] = cin[22] = cin[23] = 0.f;
            const float wgt = live ? a.weights[m] : 0.f;
            // d loss / d colour pre-activation (scaled): colour = sigmoid(z2)
            float dz2[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                const float c = a.pts_color[3 * m + o];
                dz2[o] = live ? (wgt * gi[o]) * (c * (1.0f - c)) * scale : 0.f;
            }
            if (live) {
#pragma unroll
                for (int q = 0; q < 21; ++q) tB[(size_t)(64 + q) * ld + m] = __float2half_rn(cin[q]);
#pragma unroll
                for (int o = 0; o < 3; ++o) tA[(size_t)(128 + o) * ld + m] = __float2half_rn(dz2[o]);
            }
            // layer 0 forward
            write_cin_row(g, cin);
            group_mma_round(g, [&] { issue_k32_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C0_HI, g.b_s + HB_C0_LO); });
            uint64_t mask1 = 0ull, mask2 = 0ull;
#pragma unroll 1
            for (int qtr = 0; qtr < 4; ++qtr) {
                float acc[16];
                tc05::tmem_ld16(g.tmem + qtr * 16, acc);
                uint32_t bits = 0u;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    bits |= (acc[jj] > 0.f ? 1u : 0u) << jj;
                    acc[jj] = fmaxf(acc[jj], 0.f);
                    if (live) tB[(size_t)(qtr * 16 + jj) * ld + m] = __float2half_rn(acc[jj]);
                }
                mask1 |= (uint64_t)bits << (16 * qtr);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 pk;
                    pk.x = tc05::pack_f16x2(acc[8 * half + 0], acc[8 * half + 1]); pk.y = tc05::pack_f16x2(acc[8 * half + 2], acc[8 * half + 3]);
                    pk.z = tc05::pack_f16x2(acc[8 * half + 4], acc[8 * half + 5]); pk.w = tc05::pack_f16x2(acc[8 * half + 6], acc[8 * half + 7]);
                    *reinterpret_cast<uint4*>(g.a + (2 * qtr + half) * 2048 + g.row * 16) = pk;
                }
            }
            // layer 1 forward -> relu mask, h2; d loss / d a2 = (C2^T dz2) * [a2 > 0] -> A tile of the first data-gradient MMA
            group_mma_round(g, [&] { issue_k64_x2<64>(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C1_HI, g.b_s + HB_C1_LO, 1024u); });
#pragma unroll 1
            for (int qtr = 0; qtr < 4; ++qtr) {
                float acc[16];
                tc05::tmem_ld16(g.tmem + qtr * 16, acc);
                const float4* __restrict__ c2 = reinterpret_cast<const float4*>(c2t + qtr * 16 * 4);
                uint32_t bits = 0u;
                float da[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const bool on = acc[jj] > 0.f;
                    bits |= (on ? 1u : 0u) << jj;
                    const float4 w = c2[jj];
                    const float dh = fmaf(w.x, dz2[0], fmaf(w.y, dz2[1], w.z * dz2[2]));
                    da[jj] = on ? dh : 0.f;
                    if (live) {
                        tB[(size_t)(96 + qtr * 16 + jj) * ld + m] = __float2half_rn(fmaxf(acc[jj], 0.f));
                        tA[(size_t)(qtr * 16 + jj) * ld + m] = __float2half_rn(da[jj]);
                    }
                }
                mask2 |= (uint64_t)bits << (16 * qtr);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 pk;
                    pk.x = tc05::pack_f16x2(da[8 * half + 0], da[8 * half + 1]); pk.y = tc05::pack_f16x2(da[8 * half + 2], da[8 * half + 3]);
                    pk.z = tc05::pack_f16x2(da[8 * half + 4], da[8 * half + 5]); pk.w = tc05::pack_f16x2(da[8 * half + 6], da[8 * half + 7]);
                    *reinterpret_cast<uint4*>(g.a + (2 * qtr + half) * 2048 + g.row * 16) = pk;
                }
            }
            (void)mask2;
            // d loss / d h1 = da2 C1 -> d loss / d a1 = (.) * [a1 > 0]
            group_mma_round(g, [&] { issue_k64_x2<64>(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C1T_HI, g.b_s + HB_C1T_LO, 1024u); });
#pragma unroll 1
            for (int qtr = 0; qtr < 4; ++qtr) {
                float acc[16];
                tc05::tmem_ld16(g.tmem + qtr * 16, acc);
                const uint32_t bits = (uint32_t)(mask1 >> (16 * qtr)) & 0xFFFFu;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    acc[jj] = ((bits >> jj) & 1u) ? acc[jj] : 0.f;
                    if (live) tA[(size_t)(64 + qtr * 16 + jj) * ld + m] = __float2half_rn(acc[jj]);
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint4 pk;
                    pk.x = tc05::pack_f16x2(acc[8 * half + 0], acc[8 * half + 1]); pk.y = tc05::pack_f16x2(acc[8 * half + 2], acc[8 * half + 3]);
                    pk.z = tc05::pack_f16x2(acc[8 * half + 4], acc[8 * half + 5]); pk.w = tc05::pack_f16x2(acc[8 * half + 6], acc[8 * half + 7]);
                    *reinterpret_cast<uint4*>(g.a + (2 * qtr + half) * 2048 + g.row * 16) = pk;
                }
            }
            // d loss / d cin = da1 C0 (N = 32)
            group_mma_round(g, [&] { issue_k64_x2<32>(g.tmem & 0xFFFFu, g.a_s, g.b_s + HB_C0T_HI, g.b_s + HB_C0T_LO, 512u); });
            float dcin[32];
            {
                float acc[16];
                tc05::tmem_ld16(g.tmem, acc);
#pragma unroll
                for (int q = 0; q < 16; ++q) dcin[q] = acc[q] * inv_scale;
                tc05::tmem_ld16(g.tmem + 16, acc);
#pragma unroll
                for (int q = 0; q < 16; ++q) dcin[16 + q] = acc[q] * inv_scale;
            }
            // ---- alpha chain (models/instant_nsr.py:226-250) ----
            const float da_raw = (s.ratio >= 0.0f && s.ratio <= 1.0f) ? dalpha[b] : 0.f;       // clip passes the gradient inside [0, 1]
            const float den = s.c0 + 1e-5f, num = s.c0 - s.c1 + 1e-5f;
            const float dnum = da_raw / den, dden = -da_raw * num / (den * den);
            const float dc0 = dnum + dden, dc1 = -dnum;
            const float dx0 = dc0 * (s.c0 * (1.0f - s.c0)), dx1 = dc1 * (s.c1 * (1.0f - s.c1));
            const float dsdf = (dx0 + dx1) * inv_s;
            const float dhalf = (dx1 - dx0) * inv_s;
            if (live) ds_acc += dx0 * (s.sdf - s.hs) + dx1 * (s.sdf + s.hs);
            const float dit = dhalf * s.gap * 0.5f;
            const float car = a.cos_anneal_ratio;
            const float u = -s.cosv * 0.5f + 0.5f, v = -s.cosv;
            const float spu = u * 100.0f > 20.0f ? 1.0f : sigmoidf(u * 100.0f), spv = v * 100.0f > 20.0f ? 1.0f : sigmoidf(v * 100.0f);
            const float dcos = dit * (0.5f * (1.0f - car) * spu + car * spv);
            // ---- normal: from the colour MLP input, the normal map, the cosine ----
            float dn[3];
            dn[0] = dcin[3] + wgt * gnm[0] + r.dx * dcos;
            dn[1] = dcin[4] + wgt * gnm[1] + r.dy * dcos;
            dn[2] = dcin[5] + wgt * gnm[2] + r.dz * dcos;
            const float inv = 1e-5f + s.gn;
            const float dot = dn[0] * s.gx + dn[1] * s.gy + dn[2] * s.gz;
            float dgn = -dot / (inv * inv);                                                      // through 1 / (1e-5 + |g|)
            const float pn = sqrtf(s.px * s.px + s.py * s.py + s.pz * s.pz);
            if (pn < 1.2f) dgn += g_eik * 2.0f * (s.gn - 1.0f) / eik_den;                         // eikonal term
            const float rg = s.gn > 0.f ? dgn / s.gn : 0.f;
            const float dg0 = dn[0] / inv + rg * s.gx, dg1 = dn[1] / inv + rg * s.gy, dg2 = dn[2] / inv + rg * s.gz;
            if (live) {
                const float hf = 0.5f / eps;
                gr.g_fd[m] = hf * dg0; gr.g_fd[(size_t)M + m] = -hf * dg0;
                gr.g_fd[2 * (size_t)M + m] = hf * dg1; gr.g_fd[3 * (size_t)M + m] = -hf * dg1;
                gr.g_fd[4 * (size_t)M + m] = hf * dg2; gr.g_fd[5 * (size_t)M + m] = -hf * dg2;
                float4* gc = reinterpret_cast<float4*>(gr.g_centre + 16 * m);
                gc[0] = make_float4(dsdf, dcin[6], dcin[7], dcin[8]);
                gc[1] = make_float4(dcin[9], dcin[10], dcin[11], dcin[12]);
                gc[2] = make_float4(dcin[13], dcin[14], dcin[15], dcin[16]);
                gc[3] = make_float4(dcin[17], dcin[18], dcin[19], dcin[20]);
                b1_acc[0] += dsdf;               // the six neighbours' +hf*dg and -hf*dg cancel in the bias of the signed distance
#pragma unroll
                for (int q = 1; q < 16; ++q) b1_acc[q] += dcin[5 + q];
            }
        }
    }
    // d loss / d variance = 10 * inv_s * sum(d loss / d inv_s)
    ds_acc = warp_sum(ds_acc);
    if (lane == 0 && gr.g_variance && s_live && ds_acc != 0.f) atomicAdd(gr.g_variance, ds_acc * 10.0f * inv_s);
    if (gr.g_b1) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float v = warp_sum(b1_acc[q]);
            if (lane == 0 && v != 0.f) atomicAdd(gr.g_b1 + q, v);
        }
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64 * kGroupsH>(tmem_base);
}

// Section points of the sorted depths (models/instant_nsr.py:186-206): mid-point of every interval (the last sample keeps
// its own depth), clamped to the bound.  One thread per sample.
__global__ void __launch_bounds__(256) section_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                             const float* __restrict__ z, uint32_t n, uint32_t T, float bound, float* __restrict__ P) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * T) return;
    const uint32_t ray = (uint32_t)(i / T), k = (uint32_t)(i - (size_t)ray * T);
    const float zk = z[i];
    const float zm = k < T - 1 ? zk + 0.5f * (z[i + 1] - zk) : zk;
    Ray r;
    r.ox = rays_o[3 * ray]; r.oy = rays_o[3 * ray + 1]; r.oz = rays_o[3 * ray + 2];
    r.dx = rays_d[3 * ray]; r.dy = rays_d[3 * ray + 1]; r.dz = rays_d[3 * ray + 2];
    float x, y, zz;
    ray_point(r, zm, x, y, zz);
    P[3 * i] = clampf(x, -bound, bound); P[3 * i + 1] = clampf(y, -bound, bound); P[3 * i + 2] = clampf(zz, -bound, bound);
}

// Stage kernels of the warped (render_can=False) path, whose points go through the SMPL warp between the stages
// (models/instant_nsr.py:155-172, :461-475):
//   ray_points_kernel   z [n,T] (or, z == NULL, the coarse depths near + (far - near) * linspace(0, 1, T) written to z_out) ->
//                       points o + d z, optionally clamped to +-bound
//   merge_gather_kernel sdf [n,T] and s_new [n,16] gathered by the merge permutation `order` [n,T+16] (cat_z_vals :466-470)
__global__ void __launch_bounds__(256) ray_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
                                                         const float* __restrict__ near_far, uint32_t n, uint32_t T, float bound,
                                                         float* __restrict__ z_out, float* __restrict__ P) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * T) return;
    const uint32_t ray = (uint32_t)(i / T), k = (uint32_t)(i - (size_t)ray * T);
    float zk;
    if (z) {
        zk = z[i];
    } else {
        const float near = near_far[2 * ray], far = near_far[2 * ray + 1];
        zk = near + (far - near) * linspace01((int)k, (int)T);
        z_out[i] = zk;
    }
    Ray r;
    r.ox = rays_o[3 * ray]; r.oy = rays_o[3 * ray + 1]; r.oz = rays_o[3 * ray + 2];
    r.dx = rays_d[3 * ray]; r.dy = rays_d[3 * ray + 1]; r.dz = rays_d[3 * ray + 2];
    float x, y, zz;
    ray_point(r, zk, x, y, zz);
    if (bound > 0.f) { x = clampf(x, -bound, bound); y = clampf(y, -bound, bound); zz = clampf(zz, -bound, bound); }
    P[3 * i] = x; P[3 * i + 1] = y; P[3 * i + 2] = zz;
}
__global__ void __launch_bounds__(256) merge_gather_kernel(const float* __restrict__ sdf, const float* __restrict__ s_new, const int32_t* __restrict__ order,
                                                           uint32_t n, uint32_t T, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t To = T + 16;
    if (i >= (size_t)n * To) return;
    const uint32_t ray = (uint32_t)(i / To);
    const int32_t src = order[i];
    out[i] = src < (int32_t)T ? sdf[(size_t)ray * T + src] : s_new[(size_t)ray * 16 + (src - (int32_t)T)];
}

// x[i] = clamp(x[i], -bound, bound) in place (points after the warp, :172) ; sdf[b] = out16[b][0] (the signed distance column)
__global__ void __launch_bounds__(256) clamp_kernel(float* __restrict__ x, size_t n, float bound) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = clampf(x[i], -bound, bound);
}
__global__ void __launch_bounds__(256) take_sdf_kernel(const float* __restrict__ out16, size_t B, float* __restrict__ sdf) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (size_t)gridDim.x * blockDim.x) sdf[i] = out16[16 * i];
}

// Weight-norm backward for one layer: W = g * v / |v|_row  ->  dg = (dW . v) / |v|,  dv = g / |v| * (dW - (dW . v) v / |v|^2).
// One warp per row; gradients are ADDED to dv / dg (the flat gradient buffer of the optimiser).
struct WnLayer { const float* dW; const float* v; const float* g; float* dv; float* dg; int rows, cols, ldw; const float* scale; float* db; int db_col; };
struct WnArgs { WnLayer layer[5]; };
__global__ void __launch_bounds__(256) weight_norm_backward_kernel(const WnArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int row = warp, li = 0;
    while (li < 5 && row >= a.layer[li].rows) { row -= a.layer[li].rows; ++li; }
    if (li >= 5) return;
    const WnLayer L = a.layer[li];
    const float sc = L.scale ? 1.0f / L.scale[0] : 1.0f;            // dW was accumulated times a power-of-two operand scale
    float vv = 0.f, dv_dot = 0.f;
    for (int c = lane; c < L.cols; c += 32) {
        const float v = L.v[row * L.cols + c], d = L.dW[row * L.ldw + c] * sc;
        vv = fmaf(v, v, vv); dv_dot = fmaf(d, v, dv_dot);
    }
    vv = warp_sum(vv); dv_dot = warp_sum(dv_dot);
    const float nrm = sqrtf(vv), gg = L.g[row];
    if (lane == 0) {
        L.dg[row] += dv_dot / nrm;
        if (L.db) L.db[row] += L.dW[row * L.ldw + L.db_col] * sc;   // bias gradient kept in a column of the same accumulator
    }
    for (int c = lane; c < L.cols; c += 32) {
        const float v = L.v[row * L.cols + c], d = L.dW[row * L.ldw + c] * sc;
        L.dv[row * L.cols + c] += gg / nrm * (d - dv_dot * v / vv);
    }
}

// (s_d, s_g) for ac_nsr_sdf_backward_stencil: powers of two with s_g <= 30000 / gmax and s_d <= 30000 / (gmax * c1), gmax = the
// largest |upstream gradient| over g_centre and g_fd, c1 = max_j sum_o |W1[o][j]| (see the header of that entry point).
__global__ void __launch_bounds__(1024) sdf_backward_scales_kernel(const float* __restrict__ g_centre, size_t n_centre, const float* __restrict__ g_fd,
                                                                   size_t n_fd, const float* __restrict__ blob, float* __restrict__ scales) {
    __shared__ float sm[32];
    float mx = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_centre; i += (size_t)gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(g_centre[i]));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_fd; i += (size_t)gridDim.x * blockDim.x) mx = fmaxf(mx, fabsf(g_fd[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        mx = sm[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(scales + 2), __float_as_uint(mx));     // non-negative floats order like uints
    }
}
__global__ void __launch_bounds__(64) sdf_backward_scales_finish_kernel(const float* __restrict__ blob, float* __restrict__ scales) {
    float c = 0.f;
    for (int o = 0; o < 16; ++o) c += fabsf(blob[OFF_W1T + threadIdx.x * 16 + o]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
    __shared__ float s2[2];
    if ((threadIdx.x & 31) == 0) s2[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float c1 = fmaxf(fmaxf(s2[0], s2[1]), 1e-30f), gmax = fmaxf(scales[2], 1e-30f);
        scales[1] = exp2f(fminf(fmaxf(floorf(log2f(30000.0f / gmax)), -100.f), 100.f));
        scales[0] = exp2f(fminf(fmaxf(floorf(log2f(30000.0f / (gmax * c1))), -100.f), 100.f));
    }
}

__global__ void __launch_bounds__(256) fill_uniform_kernel(float* __restrict__ out, size_t n, uint64_t seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = seed * 0xD1342543DE82EF95ull + i;
        z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        out[i] = (float)(z >> 40) * (1.0f / 16777216.0f);            // 24 random bits -> [0, 1)
    }
}

int check_shade(const ac_nsr_model* m, const ac_nsr_shade_args* a) {
    if (!m || !m->mlp_blob || !m->variance || !a) return AC_E_INVALID_ARG;
    if (!a->rays_o || !a->rays_d || !a->z_vals || !a->points || !a->centre || !a->fd) return AC_E_INVALID_ARG;
    if (!a->rgb || !a->depth || !a->weight_sum || !a->normal || !a->eik_partial || !a->weights || !a->pts_color || !a->pts_alpha) return AC_E_INVALID_ARG;
    if (a->n_samples < 2 || a->n_samples > (uint32_t)kMaxTH || a->num_steps < 1 || !(a->eps > 0.f)) return AC_E_INVALID_ARG;
    return AC_OK;
}

uint32_t shade_grid(uint32_t n_rays) {
    const uint32_t n_quads = (n_rays + 3) / 4, sms = (uint32_t)acb::sm_count();
    return n_quads < sms ? n_quads : sms;
}

}  // namespace

extern "C" {

int ac_nsr_shade_forward(const ac_nsr_model* m, const ac_nsr_shade_args* a, float* eik_out, void* stream) {
    if (int rc = check_shade(m, a)) return rc;
    if (!eik_out) return AC_E_INVALID_ARG;
    if (a->n_rays == 0) return AC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ShadeParams p; p.a = *a; p.blob = m->mlp_blob; p.variance = m->variance;
    ACB_SET_MAX_SMEM(shade_forward_kernel, SH_TOTAL);
    shade_forward_kernel<<<shade_grid(a->n_rays), kWarpsH * 32, SH_TOTAL, st>>>(p);
    if (int rc = acb::launched()) return rc;
    shade_eik_reduce_kernel<<<1, 1024, 0, st>>>(a->eik_partial, a->n_rays, eik_out);
    return acb::launched();
}

int ac_nsr_shade_backward(const ac_nsr_model* m, const ac_nsr_shade_args* a, const ac_nsr_shade_grads* g, void* stream) {
    if (int rc = check_shade(m, a)) return rc;
    if (!g || !g->g_rgb || !g->eik_out || !g->scale || !g->g_centre || !g->g_fd || !g->terms_a || !g->terms_b) return AC_E_INVALID_ARG;
    if (g->terms_ld < (uint64_t)a->n_rays * a->n_samples || (g->terms_ld & 7u)) return AC_E_INVALID_ARG;
    if (a->n_rays == 0) return AC_OK;
    ShadeParams p; p.a = *a; p.blob = m->mlp_blob; p.variance = m->variance;
    ACB_SET_MAX_SMEM(shade_backward_kernel, SH_TOTAL);
    shade_backward_kernel<<<shade_grid(a->n_rays), kWarpsH * 32, SH_TOTAL, (cudaStream_t)stream>>>(p, *g);
    return acb::launched();
}

int ac_nsr_sdf_backward_scales(const ac_nsr_model* m, const float* g_centre, const float* g_fd, uint32_t M, float* scales, void* stream) {
    if (!m || !m->mlp_blob || !g_centre || !g_fd || !scales) return AC_E_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scales, 0, 3 * sizeof(float), st) != cudaSuccess) return acb::cuda_fail();
    if (M > 0) {
        sdf_backward_scales_kernel<<<acb::sm_count(), 1024, 0, st>>>(g_centre, (size_t)M * 16, g_fd, (size_t)M * 6, m->mlp_blob, scales);
        if (int rc = acb::launched()) return rc;
    }
    sdf_backward_scales_finish_kernel<<<1, 64, 0, st>>>(m->mlp_blob, scales);
    return acb::launched();
}

int ac_fill_uniform(float* out, uint64_t n, uint64_t seed, void* stream) {
    if (!out) return AC_E_INVALID_ARG;
    if (n == 0) return AC_OK;
    uint64_t want = (n + 255) / 256, cap = (uint64_t)acb::sm_count() * 8;
    fill_uniform_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(out, (size_t)n, seed);
    return acb::launched();
}

int ac_zero(void* p, uint64_t bytes, void* stream) {
    if (!p) return AC_E_INVALID_ARG;
    return cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream) == cudaSuccess ? AC_OK : acb::cuda_fail();
}

int ac_absmax_scale(const float* x, uint32_t n, float target, float* scale_out, void* stream) {
    if (!x || !scale_out || !(target > 0.f)) return AC_E_INVALID_ARG;
    absmax_scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, target, scale_out);
    return acb::launched();
}

int ac_nsr_section_points(const float* rays_o, const float* rays_d, const float* z_vals, uint32_t n_rays, uint32_t n_samples, float bound,
                          float* points, void* stream) {
    if (!rays_o || !rays_d || !z_vals || !points || n_samples < 1) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    const size_t total = (size_t)n_rays * n_samples;
    section_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z_vals, n_rays, n_samples, bound, points);
    return acb::launched();
}

int ac_nsr_ray_points(const float* rays_o, const float* rays_d, const float* z, const float* near_far, uint32_t n_rays, uint32_t n_samples, float bound,
                      float* z_out, float* points, void* stream) {
    if (!rays_o || !rays_d || !points || n_samples < 2 || (!z && (!near_far || !z_out))) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    const size_t total = (size_t)n_rays * n_samples;
    ray_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, near_far, n_rays, n_samples, bound, z_out, points);
    return acb::launched();
}

int ac_clamp_inplace(float* x, uint64_t n, float bound, void* stream) {
    if (!x || !(bound > 0.f)) return AC_E_INVALID_ARG;
    if (n == 0) return AC_OK;
    uint64_t want = (n + 255) / 256, cap = (uint64_t)acb::sm_count() * 16;
    clamp_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, bound);
    return acb::launched();
}

int ac_nsr_take_sdf(const float* out16, uint64_t B, float* sdf, void* stream) {
    if (!out16 || !sdf) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    uint64_t want = (B + 255) / 256, cap = (uint64_t)acb::sm_count() * 16;
    take_sdf_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(out16, (size_t)B, sdf);
    return acb::launched();
}

int ac_nsr_merge_gather(const float* sdf, const float* s_new, const int32_t* order, uint32_t n_rays, uint32_t T, float* out, void* stream) {
    if (!sdf || !s_new || !order || !out || T < 1) return AC_E_INVALID_ARG;
    if (n_rays == 0) return AC_OK;
    const size_t total = (size_t)n_rays * (T + 16);
    merge_gather_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sdf, s_new, order, n_rays, T, out);
    return acb::launched();
}

int ac_nsr_weight_norm_backward(const ac_weight_norm_layer* layers, uint32_t n_layers, void* stream) {
    if (!layers || n_layers == 0 || n_layers > 5) return AC_E_INVALID_ARG;
    WnArgs a;
    int rows = 0;
    for (uint32_t i = 0; i < 5; ++i) {
        if (i < n_layers) {
            const ac_weight_norm_layer& L = layers[i];
            if (!L.dW || !L.v || !L.g || !L.dv || !L.dg || L.rows <= 0 || L.cols <= 0 || L.ldw < L.cols) return AC_E_INVALID_ARG;
            if (L.db && (L.db_col < 0 || L.db_col >= L.ldw)) return AC_E_INVALID_ARG;
            a.layer[i] = {L.dW, L.v, L.g, L.dv, L.dg, L.rows, L.cols, L.ldw, L.scale, L.db, L.db_col};
            rows += L.rows;
        } else {
            a.layer[i] = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, nullptr, nullptr, 0};
        }
    }
    weight_norm_backward_kernel<<<(rows * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    return acb::launched();
}

}  // extern "C"
