// nsr_render_st.cuh -- the fused Instant-NSR render kernel with a group-cooperative finite-difference stencil
// (included by nsr_render_tc.cu only; uses its Group / MMA / epilogue helpers).
//
// Why: ncu on the round-1 kernel (one point per lane) showed the L1 data pipe at 73-77 % with 277 M global-load
// requests per 256x256 frame = exactly one 8-byte gather per (point, level, corner): the six +-0.005 neighbours of a
// section point (models/instant_nsr.py:683-704) re-fetched the eight corners of the SAME grid cell on every level
// coarser than ~12.  Here a lane owns a SAMPLE and evaluates all 7 stencil points of a level from one set of
// registers:
//
//   * the centre cell's 8 corners are loaded once per (sample, level);
//   * a neighbour along +-axis that stays in the cell re-blends those registers with its own weights;
//   * a neighbour that steps into the adjacent cell (it can only step ONE cell while eps*scale/(2*bound) < 1, and
//     only in the direction of its sign) loads the 4 corners of the one new face; the blend runs over the three
//     faces (old-low, old-high, new) in the reference's corner order with a ZERO weight on the face that does not
//     belong to the neighbour's cell: fma(0, v, r) == r, so every feature is bit-identical to an independent
//     8-corner evaluation (hashencoder.cu:121-172);
//   * levels fine enough for the stencil to span several cells (12..15 for eps = 0.005, bound = 1.6) gather all 8
//     corners per neighbour as before.
//
// Work split: a GROUP of 4 warps processes 32 consecutive samples of ONE ray per pass.  Warp w encodes levels
// {w, w+4, w+8, w+12} of all 7 points (one dense-coarse, one mid, one shared-hashed and one fine level each: the warps
// stay balanced and the level kind is warp-uniform), i.e. one 16-byte K chunk of every A row; the W0 tile used for
// these passes has its K columns permuted to match.  The 7 x 32 rows fill two 128-row MMA tiles (slot = 32 rows =
// one stencil point; the centre's slot rotates with the ray so that the warp that owns the ray gets the 16-output
// tail and the other three get two 1-output tails each -- tcgen05.ld restricts warp w to TMEM lanes 32w..32w+31).
// After the four rays of the group have had their pass, every warp shades its own ray's 32 samples exactly as the
// round-1 kernel did (colour MLP on tcgen05, NeuS alpha, shuffle-scan compositing).
//
// Coarse sampling and the importance rounds are the round-1 code (warp owns ray, lane = point, natural K order):
// sample depths are bit-identical to the round-1 kernel.
#pragma once

namespace {

#ifndef AC_MAGIC_FLOOR
#define AC_MAGIC_FLOOR 1     // floor via add-round-down of 2^23 (two FADDs + one LOP) instead of FRND + F2I on the XU pipe
#endif

constexpr int kGroupsS = 4;                     // 4-warp groups per CTA: 512 threads, <= 128 registers
constexpr int kWarpsS = 4 * kGroupsS;

// dynamic shared memory map (bytes)
constexpr size_t SS_LEVELS = 0;
constexpr size_t SS_B = (SS_LEVELS + kLevels * sizeof(LevelMeta) + 127) / 128 * 128;
constexpr uint32_t B_W0P_HI = B_BYTES, B_W0P_LO = B_BYTES + 4096, BS_BYTES = B_BYTES + 8192;   // W0 with permuted K columns
constexpr size_t SS_A = SS_B + BS_BYTES;                                    // per group 32 KB: two tiles of (hi 8 KB | lo 8 KB)
constexpr size_t SS_F = SS_A + (size_t)kGroupsS * 32768;                    // per group 8 KB: the centre's 15 geometry features as colour-MLP
                                                                            // A chunks (hi c0 c1 | lo c0 c1), written by the 16-output tail
constexpr size_t SS_ROWS = SS_F + (size_t)kGroupsS * 8192;                  // per warp: depths, sdf, scratch (3 x 128 floats)
constexpr size_t SS_RAYS = SS_ROWS + (size_t)kWarpsS * 3 * kMaxT * 4;       // per warp: origin, direction, near, span (8 floats)
constexpr size_t SS_BARS = SS_RAYS + (size_t)kWarpsS * 32;
constexpr size_t SS_TOTAL = SS_BARS + kGroupsS * 8 + 16;
static_assert(SS_TOTAL <= 227 * 1024, "shared memory budget");

// Packed fp32 FMA (Blackwell FFMA2): r = w * v + r on both halves, each IEEE-rounded exactly like fmaf.
__device__ __forceinline__ void ffma2(float2& r, float w, float2 v) {
    unsigned long long rr = *reinterpret_cast<unsigned long long*>(&r);
    const float2 ww = make_float2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(rr)
        : "l"(*reinterpret_cast<const unsigned long long*>(&ww)), "l"(*reinterpret_cast<const unsigned long long*>(&v)));
    r = *reinterpret_cast<float2*>(&rr);
}

// floorf(p) and its integer value for 0 <= p < 2^22 (cell positions: p = u * scale + 0.5 with u in [0,1]).
__device__ __forceinline__ void floor_pos(float p, float& fl, uint32_t& ci) {
#if AC_MAGIC_FLOOR
    const float t = __fadd_rd(p, 8388608.0f);           // 2^23 + floor(p), exact
    fl = t - 8388608.0f;
    ci = __float_as_uint(t) & 0x7FFFFFu;
#else
    fl = floorf(p);
    ci = (uint32_t)fl;
#endif
}

struct Cell {            // the centre point's cell on one level
    float q[3], f[3];    // 1 - frac, frac
    uint32_t c[3];       // cell coordinates
    uint32_t t[3];       // slot terms of the low corner: x, y * m1, z * m2  (dense: strides, hashed: primes)
    uint32_t mul[3];     // 1, m1, m2
    uint32_t mask;
};

template <bool HASHED>
__device__ __forceinline__ uint32_t slot_of(uint32_t x, uint32_t y, uint32_t z, uint32_t mask) {
    return HASHED ? ((x ^ y ^ z) & mask) : (x + y + z);
}

// Stencil neighbour along axis A (sign PLUS) on a level where it can leave the centre cell by at most one cell.
// C = the centre cell's corners (index = xbit + 2 ybit + 4 zbit); un = the neighbour's normalised coordinate on A.
template <bool HASHED, int A, bool PLUS>
__device__ __forceinline__ float2 neighbour_shared(const float2* __restrict__ t, const Cell& ce, const float2 (&C)[8], float un, float scale) {
    const float pa = fmaf(un, scale, 0.5f);
    float fl; uint32_t ni;
    floor_pos(pa, fl, ni);
    const float fa = pa - fl, qa = 1.0f - fa;
    const bool cross = ni != ce.c[A];
    // the one face the centre cell does not have: cell coordinate c+2 (PLUS) or c-1 along A
    const uint32_t ta = PLUS ? ce.t[A] + 2u * ce.mul[A] : ce.t[A] - ce.mul[A];
    constexpr int B0 = A == 0 ? 1 : 0, B1 = A == 2 ? 1 : 2;         // the two other axes, lower first
    float2 L[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t term[3];
        term[A] = ta;
        term[B0] = ce.t[B0] + ((k & 1) ? ce.mul[B0] : 0u);
        term[B1] = ce.t[B1] + ((k & 2) ? ce.mul[B1] : 0u);
        L[k] = make_float2(0.f, 0.f);
        if (cross) L[k] = __ldg(t + slot_of<HASHED>(term[0], term[1], term[2], ce.mask));
    }
    // weights of the three faces along A, in ascending cell order: PLUS (old-low, old-high, new), MINUS (new, old-low, old-high)
    float w3[3];
    if (PLUS) { w3[0] = cross ? 0.f : qa; w3[1] = cross ? qa : fa; w3[2] = cross ? fa : 0.f; }
    else      { w3[0] = cross ? qa : 0.f; w3[1] = cross ? fa : qa; w3[2] = cross ? 0.f : fa; }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int iz = 0; iz < (A == 2 ? 3 : 2); ++iz)
#pragma unroll
        for (int iy = 0; iy < (A == 1 ? 3 : 2); ++iy)
#pragma unroll
            for (int ix = 0; ix < (A == 0 ? 3 : 2); ++ix) {
                const int i3[3] = {ix, iy, iz};
                const float wx = A == 0 ? w3[ix] : (ix ? ce.f[0] : ce.q[0]);
                const float wy = A == 1 ? w3[iy] : (iy ? ce.f[1] : ce.q[1]);
                const float wz = A == 2 ? w3[iz] : (iz ? ce.f[2] : ce.q[2]);
                const float w = (wx * wy) * wz;                        // the reference's product order (hashencoder.cu:141-153)
                const int pa3 = i3[A];
                const bool is_new = PLUS ? pa3 == 2 : pa3 == 0;
                const int abit = PLUS ? pa3 : pa3 - 1;
                float2 v;
                if (is_new) v = L[i3[B0] + 2 * i3[B1]];
                else {
                    int bits[3] = {ix, iy, iz};
                    bits[A] = abit;
                    v = C[bits[0] + 2 * bits[1] + 4 * bits[2]];
                }
                ffma2(r, w, v);
            }
    return r;
}

// One level of the stencil: out[0] = centre, out[1..6] = +x, -x, +y, -y, +z, -z.  uc = centre (normalised), un[n] = the
// moved coordinate of neighbour n (normalised).  SHARED: neighbours can leave the centre cell by at most one cell.
template <bool HASHED, bool SHARED>
__device__ __forceinline__ void stencil_level(const float2* __restrict__ table, const LevelMeta m, const float (&uc)[3], const float (&un)[6],
                                              float2 (&out)[7]) {
    if constexpr (!SHARED) {
        out[0] = grid_level_3d_k<HASHED>(table, m, uc[0], uc[1], uc[2]);
#pragma unroll
        for (int n = 0; n < 6; ++n) {
            const int a = n >> 1;
            out[1 + n] = grid_level_3d_k<HASHED>(table, m, a == 0 ? un[n] : uc[0], a == 1 ? un[n] : uc[1], a == 2 ? un[n] : uc[2]);
        }
    } else {
        Cell ce;
        const float2* __restrict__ t = table + m.offset;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float p = fmaf(uc[d], m.scale, 0.5f);
            float fl;
            floor_pos(p, fl, ce.c[d]);
            ce.f[d] = p - fl;
            ce.q[d] = 1.0f - ce.f[d];
        }
        ce.mul[0] = 1u;
        ce.mul[1] = HASHED ? 2654435761u : m.res1;
        ce.mul[2] = HASHED ? 805459861u : m.res1 * m.res1;
        ce.t[0] = ce.c[0]; ce.t[1] = ce.c[1] * ce.mul[1]; ce.t[2] = ce.c[2] * ce.mul[2];
        ce.mask = m.size - 1u;
        float2 C[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            C[k] = __ldg(t + slot_of<HASHED>(ce.t[0] + (k & 1), ce.t[1] + ((k & 2) ? ce.mul[1] : 0u), ce.t[2] + ((k & 4) ? ce.mul[2] : 0u), ce.mask));
        float2 r = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = (((k & 1) ? ce.f[0] : ce.q[0]) * ((k & 2) ? ce.f[1] : ce.q[1])) * ((k & 4) ? ce.f[2] : ce.q[2]);
            ffma2(r, w, C[k]);
        }
        out[0] = r;
        out[1] = neighbour_shared<HASHED, 0, true>(t, ce, C, un[0], m.scale);
        out[2] = neighbour_shared<HASHED, 0, false>(t, ce, C, un[1], m.scale);
        out[3] = neighbour_shared<HASHED, 1, true>(t, ce, C, un[2], m.scale);
        out[4] = neighbour_shared<HASHED, 1, false>(t, ce, C, un[3], m.scale);
        out[5] = neighbour_shared<HASHED, 2, true>(t, ce, C, un[4], m.scale);
        out[6] = neighbour_shared<HASHED, 2, false>(t, ce, C, un[5], m.scale);
    }
}

// Any level kind (cold path: non-power-of-two hashed levels, huge resolutions): seven independent evaluations.
__device__ __noinline__ void stencil_level_generic(const float2* __restrict__ table, const LevelMeta* __restrict__ lvp, const float (&uc)[3],
                                                   const float (&un)[6], float2 (&out)[7]) {
    const LevelMeta m = *lvp;
    out[0] = grid_level_3d(table, m, uc[0], uc[1], uc[2]);
#pragma unroll 1
    for (int n = 0; n < 6; ++n) {
        const int a = n >> 1;
        out[1 + n] = grid_level_3d(table, m, a == 0 ? un[n] : uc[0], a == 1 ? un[n] : uc[1], a == 2 ? un[n] : uc[2]);
    }
}

// Slot (0..7: tile = slot / 4, rows 32 * (slot % 4) ..) of stencil point pt (0 = centre, 1..6 = neighbours) in the pass
// of the group's ray j: the centre sits in slot j of tile 0, slot j of tile 1 stays empty, the neighbours fill the rest
// in order.
__device__ __forceinline__ int stencil_slot(int pt, int j) {
    if (pt == 0) return j;
    const int n = pt - 1, h = n >= 3 ? 1 : 0, r = n - 3 * h;
    return 4 * h + r + (r >= j ? 1 : 0);
}

template <int SLOT>
__global__ void __launch_bounds__(kWarpsS * 32, 1) nsr_render_st_kernel(const RenderParamsTC p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    LevelMeta* lv = reinterpret_cast<LevelMeta*>(smem + SS_LEVELS);
    unsigned char* bt = smem + SS_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SS_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kGroupsS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp >> 2, wq = warp & 3;

    // ---- one-time staging: level table, fp16 weight tiles (natural + K-permuted W0), barriers, TMEM ----
    {
        const float* blob = p.blob;
        if (threadIdx.x < kLevels) lv[threadIdx.x] = make_level_meta(p.offsets, threadIdx.x, p.S, p.H, 3);
        for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
            const int n = i >> 5, k = i & 31;
            stage_b_tile(bt + B_W0_HI, bt + B_W0_LO, n, k, __ldg(blob + OFF_W0 + n * kSdfInPad + 3 + k));
            // colour layer 0 with its inputs reordered: K 0..14 = geometry features (cin 6..20), K 16..21 = x, y, z, normal
            const int ck = k < 15 ? 6 + k : (k >= 16 && k < 22 ? k - 16 : -1);
            stage_b_tile(bt + B_C0_HI, bt + B_C0_LO, n, k, ck >= 0 ? __ldg(blob + OFF_C0 + n * kColInPad + ck) : 0.f);
            // K-permuted copy: chunk c = k / 8 holds levels c, c+4, c+8, c+12 (two features each)
            const int level = (k >> 3) + 4 * ((k & 7) >> 1);
            stage_b_tile(bt + B_W0P_HI, bt + B_W0P_LO, n, k, __ldg(blob + OFF_W0 + n * kSdfInPad + 3 + 2 * level + (k & 1)));
        }
        for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
            const int n = i >> 6, k = i & 63;
            stage_b_tile(bt + B_C1_HI, bt + B_C1_LO, n, k, __ldg(blob + OFF_C1 + n * kHidden + k));
        }
        // tile 1's empty slot is multiplied by the tensor core as well: keep it finite
        for (int i = threadIdx.x; i < kGroupsS * 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + SS_A)[i] = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x == 0) {
            for (int gI = 0; gI < kGroupsS; ++gI) tc05::mbar_init(bars + gI, 1);
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
    g.a = smem + SS_A + (size_t)group * 32768;
    g.a_s = tc05::smem_u32(g.a);
    g.b_s = tc05::smem_u32(bt);
    g.bar = bars + group;
    g.phase = 0;
    g.row = wq * 32 + lane;
    g.tmem = tmem_base + (uint32_t)group * 128u + ((uint32_t)(wq * 32) << 16);
    g.bar_id = 1 + group;
    bool std_levels = true;         // every level dense or power-of-two hashed, resolution inside the magic-floor range
    {
        bool ok = true;
        for (int l = 0; l < kLevels; ++l) {
            ok = ok && (lv[l].hashed == (l < 5 ? 0u : 1u));
            std_levels = std_levels && lv[l].hashed < 2u && lv[l].scale < 2.0e6f;
        }
        g.std_layout = ok;
    }

    float* rows = reinterpret_cast<float*>(smem + SS_ROWS);
    float* zs = rows + warp * 3 * kMaxT;             // sorted depths
    float* sdfs = zs + kMaxT;                        // their SDF
    float* ta = sdfs + kMaxT;                        // scratch (alpha, then cdf)
    float* rays = reinterpret_cast<float*>(smem + SS_RAYS);

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
    const float two_b = 2.0f * bound;
    const float shift_per_scale = eps / two_b;                 // stencil reach in cells = this * level scale

    // A full launch gives a CTA four adjacent quads per iteration (shared coarse cells in L1); a small one deals quads to
    // CTAs first so that every SM gets work (see launch_render_tc).
    const bool dense_map = n_quads >= gridDim.x * kGroupsS;
    for (uint32_t quad = dense_map ? blockIdx.x * kGroupsS + group : blockIdx.x + gridDim.x * group; quad < n_quads;
         quad += gridDim.x * kGroupsS) {
        const uint32_t ray_raw = quad * 4 + wq;
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

        // ---- render core (:186-299): publish the ray, then 32 section samples at a time ----
        if (lane < 8)
            rays[warp * 8 + lane] = lane == 0 ? r.ox : lane == 1 ? r.oy : lane == 2 ? r.oz : lane == 3 ? r.dx : lane == 4 ? r.dy : lane == 5 ? r.dz
                                  : lane == 6 ? near : span;
        tc05::named_bar_sync(g.bar_id, 128);            // depths and rays of all four warps visible to the group

        float carry = 1.0f;
        float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_nx = 0.f, acc_ny = 0.f, acc_nz = 0.f;
        float acc_w = 0.f, acc_d = 0.f, eik_num = 0.f, eik_den = 0.f;
        unsigned char* fsm = smem + SS_F + (size_t)group * 8192;           // colour A chunks 0, 1 of the group's 128 samples
        for (int k0 = 0; k0 < Ttot; k0 += 32) {
            float sdf0 = 0.f;                           // centre SDF of this warp's own sample (written in pass j == wq)

#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                // ===== stencil pass: 32 samples of the group's ray j, all four warps =====
                const float* rj = rays + (group * 4 + j) * 8;
                const float* zj = rows + (group * 4 + j) * 3 * kMaxT;
                float* fdj = rows + (group * 4 + j) * 3 * kMaxT + kMaxT;            // [6][32] in ray j's sdf / scratch rows
                float pc[3];
                {
                    const int k = min(k0 + lane, Ttot - 1);
                    if (p.a.pts_in) {
                        const uint32_t rayj = min(quad * 4 + (uint32_t)j, p.a.n_rays - 1);
                        const float* q3 = p.a.pts_in + 3 * ((size_t)rayj * Ttot + k);
                        pc[0] = q3[0]; pc[1] = q3[1]; pc[2] = q3[2];
                    } else {
                        const float zk = zj[k];
                        const float zmid = k < Ttot - 1 ? zk + 0.5f * (zj[k + 1] - zk) : zk;
                        pc[0] = rj[0] + rj[3] * zmid; pc[1] = rj[1] + rj[4] * zmid; pc[2] = rj[2] + rj[5] * zmid;   // mul, then add (ray_point)
                    }
                }
                {
                    // Every coordinate is clamped to [-bound, bound] first (:205, :691), so the normalised values lie in
                    // [0, 1] and the encoder's out-of-range branch (hashencoder.cu:94-119) cannot trigger.
                    float uc[3], un[6];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        pc[d] = clampf(pc[d], -bound, bound);
                        uc[d] = (pc[d] + bound) / two_b;
                        un[2 * d] = (clampf(pc[d] + eps, -bound, bound) + bound) / two_b;
                        un[2 * d + 1] = (clampf(pc[d] - eps, -bound, bound) + bound) / two_b;
                    }
                    // this warp's K chunk of the 7 rows: levels wq, wq+4 (first 8 bytes of the chunk), then wq+8, wq+12
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t hi[7][2], lo[7][2];
#pragma unroll 1
                        for (int i2 = 0; i2 < 2; ++i2) {
                            const int l = wq + 4 * (2 * hf + i2);
                            const LevelMeta m = lv[l];
                            float2 f7[7];
                            const bool shared = shift_per_scale * m.scale + 1e-3f < 1.0f;   // a neighbour leaves the centre cell by <= 1 cell
                            if (!std_levels) stencil_level_generic(table, lv + l, uc, un, f7);
                            else if (m.hashed == 0u) {
                                if (shared) stencil_level<false, true>(table, m, uc, un, f7);
                                else stencil_level<false, false>(table, m, uc, un, f7);
                            } else {
                                if (shared) stencil_level<true, true>(table, m, uc, un, f7);
                                else stencil_level<true, false>(table, m, uc, un, f7);
                            }
#pragma unroll
                            for (int pt = 0; pt < 7; ++pt) {
                                uint32_t h, lw;
                                tc05::split_f16x2(f7[pt].x, f7[pt].y, h, lw);
                                if (i2 == 0) { hi[pt][0] = h; lo[pt][0] = lw; }      // rolled loop: select, do not index registers
                                else { hi[pt][1] = h; lo[pt][1] = lw; }
                            }
                        }
#pragma unroll
                        for (int pt = 0; pt < 7; ++pt) {
                            const int sl = stencil_slot(pt, j);
                            unsigned char* arow = g.a + (sl >> 2) * 16384 + wq * 2048 + ((sl & 3) * 32 + lane) * 16 + hf * 8;
                            *reinterpret_cast<uint2*>(arow) = make_uint2(hi[pt][0], hi[pt][1]);
                            *reinterpret_cast<uint2*>(arow + 8192) = make_uint2(lo[pt][0], lo[pt][1]);
                        }
                    }
                }
                group_mma_round(g, [&] {
                    issue_k32_x3(g.tmem & 0xFFFFu, g.a_s, g.b_s + B_W0P_HI, g.b_s + B_W0P_LO);
                    issue_k32_x3((g.tmem & 0xFFFFu) + 64u, g.a_s + 16384u, g.b_s + B_W0P_HI, g.b_s + B_W0P_LO);
                });
                // tails: slot wq of tile 0 and of tile 1
                if (wq == j) {
                    float o16[16];
                    sdf_tail<SLOT, true>(g.tmem, pc[0], pc[1], pc[2], o16);
                    sdf0 = o16[0];
                    // geometry features -> colour layer 0's A chunks 0 and 1 (K 0..14, K 15 = 0), own row
                    uint4 h4, l4;
                    tc05::split_f16x2(o16[1], o16[2], h4.x, l4.x); tc05::split_f16x2(o16[3], o16[4], h4.y, l4.y);
                    tc05::split_f16x2(o16[5], o16[6], h4.z, l4.z); tc05::split_f16x2(o16[7], o16[8], h4.w, l4.w);
                    *reinterpret_cast<uint4*>(fsm + g.row * 16) = h4;
                    *reinterpret_cast<uint4*>(fsm + 4096 + g.row * 16) = l4;
                    tc05::split_f16x2(o16[9], o16[10], h4.x, l4.x); tc05::split_f16x2(o16[11], o16[12], h4.y, l4.y);
                    tc05::split_f16x2(o16[13], o16[14], h4.z, l4.z); tc05::split_f16x2(o16[15], 0.f, h4.w, l4.w);
                    *reinterpret_cast<uint4*>(fsm + 2048 + g.row * 16) = h4;
                    *reinterpret_cast<uint4*>(fsm + 4096 + 2048 + g.row * 16) = l4;
                } else {
                    const int r0 = wq - (wq > j ? 1 : 0);            // neighbour index of tile 0's slot wq; tile 1's is r0 + 3
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        const int n = r0 + 3 * h;
                        const int a = n >> 1;
                        const float e = (n & 1) ? -eps : eps;
                        float q3[3] = {pc[0], pc[1], pc[2]};
                        if (a == 0) q3[0] = clampf(pc[0] + e, -bound, bound);
                        else if (a == 1) q3[1] = clampf(pc[1] + e, -bound, bound);
                        else q3[2] = clampf(pc[2] + e, -bound, bound);
                        fdj[n * 32 + lane] = sdf_tail_scalar<SLOT>(g.tmem + 64u * h, q3[0], q3[1], q3[2]);
                    }
                }
            }
            tc05::fence_before_sync();
            tc05::named_bar_sync(g.bar_id, 128);          // all finite-difference values of the four rays are in shared memory
            tc05::fence_after_sync();

            // ===== shading: this warp's own ray, lane = sample (the round-1 kernel's arithmetic from here) =====
            const float* rw = rays + warp * 8;
            Ray rr;
            rr.ox = rw[0]; rr.oy = rw[1]; rr.oz = rw[2]; rr.dx = rw[3]; rr.dy = rw[4]; rr.dz = rw[5];
            const float near_w = rw[6], span_w = rw[7];
            const bool live = k0 + lane < Ttot;
            const int k = min(k0 + lane, Ttot - 1);
            const float zk = zs[k];
            const float delta = k < Ttot - 1 ? zs[k + 1] - zk : span_w / (float)N0;
            float px, py, pz;
            if (p.a.pts_in) {
                const float* q3 = p.a.pts_in + 3 * ((size_t)ray * Ttot + k);
                px = q3[0]; py = q3[1]; pz = q3[2];
            } else {
                ray_point(rr, k < Ttot - 1 ? zk + 0.5f * delta : zk, px, py, pz);
            }
            px = clampf(px, -bound, bound); py = clampf(py, -bound, bound); pz = clampf(pz, -bound, bound);
            float nx, ny, nz, gn;
            {
                const float* fd = sdfs;
                const float f0 = fd[lane], f1 = fd[32 + lane], f2 = fd[64 + lane], f3 = fd[96 + lane], f4 = fd[128 + lane], f5 = fd[160 + lane];
                const float gx = 0.5f * (f0 - f1) / eps, gy = 0.5f * (f2 - f3) / eps, gz = 0.5f * (f4 - f5) / eps;
                gn = sqrtf(gx * gx + gy * gy + gz * gz);
                const float inv = 1e-5f + gn;
                nx = gx / inv; ny = gy / inv; nz = gz / inv;
            }
            // colour layer 0: chunks 0, 1 (features) are already in the group's feature region; chunk 2 = (x, y, z, n, 0, 0),
            // chunk 3 = 0 go to the tile region
            {
                uint4 h4, l4;
                tc05::split_f16x2(px, py, h4.x, l4.x); tc05::split_f16x2(pz, nx, h4.y, l4.y);
                tc05::split_f16x2(ny, nz, h4.z, l4.z); h4.w = 0u; l4.w = 0u;
                *reinterpret_cast<uint4*>(g.a + 2 * 2048 + g.row * 16) = h4;
                *reinterpret_cast<uint4*>(g.a + 8192 + 2 * 2048 + g.row * 16) = l4;
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(g.a + 3 * 2048 + g.row * 16) = z4;
                *reinterpret_cast<uint4*>(g.a + 8192 + 3 * 2048 + g.row * 16) = z4;
            }
            const uint32_t f_s = tc05::smem_u32(fsm);
            group_mma_round(g, [&] {                    // its barrier also orders the fd reads before the next block's writes
                constexpr uint32_t idesc = tc05::idesc_f16(128, 64);
                const uint32_t d = g.tmem & 0xFFFFu;
                const uint32_t bh_s = g.b_s + B_C0_HI, bl_s = g.b_s + B_C0_LO;
                {   // K 0..15: feature chunks
                    const uint64_t ah = tc05::smem_desc(f_s, 2048u, 128u), al = tc05::smem_desc(f_s + 4096u, 2048u, 128u);
                    const uint64_t bh = tc05::smem_desc(bh_s, 1024u, 128u), bl = tc05::smem_desc(bl_s, 1024u, 128u);
                    tc05::mma_f16(d, ah, bh, idesc, 0u); tc05::mma_f16(d, al, bh, idesc, 1u); tc05::mma_f16(d, ah, bl, idesc, 1u);
                }
                {   // K 16..31: position / normal chunk + zero chunk
                    const uint64_t ah = tc05::smem_desc(g.a_s + 4096u, 2048u, 128u), al = tc05::smem_desc(g.a_s + 8192u + 4096u, 2048u, 128u);
                    const uint64_t bh = tc05::smem_desc(bh_s + 2048u, 1024u, 128u), bl = tc05::smem_desc(bl_s + 2048u, 1024u, 128u);
                    tc05::mma_f16(d, ah, bh, idesc, 1u); tc05::mma_f16(d, al, bh, idesc, 1u); tc05::mma_f16(d, ah, bl, idesc, 1u);
                }
            });
            float col[3];
            group_color_rest<SLOT>(g, col);
            const float cosv = rr.dx * nx + rr.dy * ny + rr.dz * nz;
            const float it = -(softplus100(-cosv * 0.5f + 0.5f) * (1.0f - car) + softplus100(-cosv) * car);
            const float hs = it * delta * 0.5f;
            const float c0 = sigmoidf((sdf0 - hs) * inv_s), c1 = sigmoidf((sdf0 + hs) * inv_s);
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
                acc_nx += w * nx; acc_ny += w * ny; acc_nz += w * nz;
                acc_w += w;
                acc_d += w * clampf((zk - near_w) / span_w, 0.0f, 1.0f);
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
        // the next quad's sampling phase overwrites the depth rows other warps of the group read during their passes
        tc05::named_bar_sync(g.bar_id, 128);
    }

    // ---- teardown ----
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace
