// raymarch_ops.cu -- the six operators of the reference's `raymarching` extension:
//   march_rays_train, composite_rays_train_forward/backward   (raymarching/src/raymarching.cu:56-391)
//   march_rays, composite_rays, compact_rays                    (raymarching/src/raymarching.cu:497-747)
// with the same buffer layouts and semantics.  NB the module is dead code in the reference (nothing imports
// it; NeRFRenderer.run_cuda is undefined, models/instant_nsr.py:362-363) -- it is provided because the
// operator API is part of the surface the north star names.  "sigmas" are already alphas (NeuS variant, :275).
//
// Structure: one occupancy-grid stepper (`GridStepper`) shared by the training and inference marchers (the
// reference spells the voxel walk out three times).  The arithmetic mirrors the reference expression by
// expression, including the FMA contractions nvcc applies to it (this TU is built --fmad=false, so they
// are explicit), because the integer outputs -- steps per ray -- depend on every rounding.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"

namespace {

constexpr float kDensityThresh = 10.0f;
constexpr int kMaxSteps = 1024;
constexpr float kSqrt3 = 1.73205080757f;
constexpr float kMinNear = 0.05f;

// PCG32 (O'Neill), seeded as the reference's pcg32(initstate, initseq) constructor does.
struct Pcg32 {
    uint64_t state, inc;
    __device__ Pcg32(uint64_t initstate, uint64_t initseq = 1u) {
        state = 0u;
        inc = (initseq << 1u) | 1u;
        next();
        state += initstate;
        next();
    }
    __device__ uint32_t next() {
        const uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    __device__ float next_float() { return __uint_as_float((next() >> 9) | 0x3f800000u) - 1.0f; }
};

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// Walks a ray through the H^3 occupancy grid of the cube [-bound, bound]^3.
struct GridStepper {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
    float bound, rbound, thresh, dt_min, dt_max, dt_gamma;
    uint32_t H;
    const float* grid;

    __device__ GridStepper(const float* o, const float* d, float bound_, uint32_t H_, const float* grid_, float mean_density)
        : ox(o[0]), oy(o[1]), oz(o[2]), dx(d[0]), dy(d[1]), dz(d[2]), bound(bound_), H(H_), grid(grid_) {
        rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
        rbound = 1 / bound;
        thresh = fminf(kDensityThresh, mean_density);
        dt_min = (2 * kSqrt3 / kMaxSteps) * bound;
        dt_max = 2 * bound / (H - 1);
        dt_gamma = bound > 1 ? 1.f / 256.f : 0.0f;
    }
    __device__ void box(float& near, float& far) const {           // :85-98
        float nx = (-bound - ox) * rdx, fx = (bound - ox) * rdx;
        if (nx > fx) { const float t = nx; nx = fx; fx = t; }
        float ny = (-bound - oy) * rdy, fy = (bound - oy) * rdy;
        if (ny > fy) { const float t = ny; ny = fy; fy = t; }
        float nz = (-bound - oz) * rdz, fz = (bound - oz) * rdz;
        if (nz > fz) { const float t = nz; nz = fz; fz = t; }
        near = fmaxf(fmaxf(nx, fmaxf(ny, nz)), kMinNear);
        far = fminf(fx, fminf(fy, fz));
    }
    __device__ float step_size(float t) const { return clampf(t * dt_gamma, dt_min, dt_max); }
    // Position at t, whether its voxel is occupied, and (if not) the parameter at which the ray leaves the voxel.
    __device__ bool probe(float t, float& x, float& y, float& z, float& t_exit) const {
        x = clampf(fmaf(t, dx, ox), -bound, bound);
        y = clampf(fmaf(t, dy, oy), -bound, bound);
        z = clampf(fmaf(t, dz, oz), -bound, bound);
        const int nx = (int)clampf((float)(0.5 * (double)fmaf(x, rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
        const int ny = (int)clampf((float)(0.5 * (double)fmaf(y, rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
        const int nz = (int)clampf((float)(0.5 * (double)fmaf(z, rbound, 1.0f) * (double)H), 0.0f, (float)(H - 1));
        if (grid[(uint32_t)nx * H * H + (uint32_t)ny * H + (uint32_t)nz] > thresh) return true;
        const float hm1 = (float)(H - 1);
        const float tx = fmaf(fmaf((nx + 0.5f + 0.5f * copysignf(1.0f, dx)) / hm1, 2.0f, -1.0f), bound, -x) * rdx;
        const float ty = fmaf(fmaf((ny + 0.5f + 0.5f * copysignf(1.0f, dy)) / hm1, 2.0f, -1.0f), bound, -y) * rdy;
        const float tz = fmaf(fmaf((nz + 0.5f + 0.5f * copysignf(1.0f, dz)) / hm1, 2.0f, -1.0f), bound, -z) * rdz;
        t_exit = t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
        return false;
    }
    __device__ float skip(float t, float t_exit) const {            // "step until next voxel"
        do { t += step_size(t); } while (t < t_exit);
        return t;
    }
};

__global__ void __launch_bounds__(256) march_rays_train_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                               const float* __restrict__ grid, float mean_density, float bound, uint32_t N,
                                                               uint32_t H, uint32_t M, float* xyzs, float* dirs, float* deltas, int* rays,
                                                               int* counter, uint32_t perturb) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const GridStepper g(rays_o + 3 * (size_t)n, rays_d + 3 * (size_t)n, bound, H, grid, mean_density);
    float near, far;
    g.box(near, far);
    float t0 = near;
    if (perturb) { Pcg32 rng((uint64_t)n); t0 = fmaf(g.dt_min, rng.next_float(), t0); }   // nvcc contracts the reference's `t0 += dt_min * u`
    // pass 1: count the occupied steps
    float t = t0, x, y, z, t_exit;
    uint32_t num_steps = 0;
    while (t < far && num_steps < (uint32_t)kMaxSteps) {
        if (g.probe(t, x, y, z, t_exit)) { ++num_steps; t += g.step_size(t); }
        else t = g.skip(t, t_exit);
    }
    // allocate a contiguous slab of samples and a ray slot (order of arrival, like the reference)
    const uint32_t point_index = (uint32_t)atomicAdd(counter, (int)num_steps);
    const uint32_t ray_index = (uint32_t)atomicAdd(counter + 1, 1);
    rays[ray_index * 3] = (int)n; rays[ray_index * 3 + 1] = (int)point_index; rays[ray_index * 3 + 2] = (int)num_steps;
    if (num_steps == 0 || point_index + num_steps >= M) return;
    // pass 2: emit
    float* px = xyzs + 3 * (size_t)point_index; float* pd = dirs + 3 * (size_t)point_index; float* pt = deltas + point_index;
    t = t0;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
        if (g.probe(t, x, y, z, t_exit)) {
            px[0] = x; px[1] = y; px[2] = z; pd[0] = g.dx; pd[1] = g.dy; pd[2] = g.dz;
            const float dt = g.step_size(t);
            t += dt;
            pt[0] = dt;
            px += 3; pd += 3; ++pt; ++step;
        } else t = g.skip(t, t_exit);
    }
}

__global__ void __launch_bounds__(256) composite_train_forward_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                      const int* __restrict__ rays, uint32_t M, uint32_t N,
                                                                      float* weights_sum, float* image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    float T = 1.0f, r = 0, g = 0, b = 0;
    if (!(num_steps == 0 || offset + num_steps >= M)) {
        for (uint32_t s = 0; s < num_steps && !(T < 1e-4f); ++s) {        // early out on spent transmittance (:273)
            const float alpha = sigmas[offset + s], w = alpha * T;
            r = fmaf(w, rgbs[3 * (size_t)(offset + s)], r);
            g = fmaf(w, rgbs[3 * (size_t)(offset + s) + 1], g);
            b = fmaf(w, rgbs[3 * (size_t)(offset + s) + 2], b);
            T *= 1.0f - alpha;
        }
        weights_sum[index] = 1.0f - T;
    } else {
        weights_sum[index] = 0;
    }
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

__global__ void __launch_bounds__(256) composite_train_backward_kernel(const float* __restrict__ grad_ws, const float* __restrict__ grad_img,
                                                                       const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                                       const float* __restrict__ deltas, const int* __restrict__ rays,
                                                                       const float* __restrict__ weights_sum, const float* __restrict__ image,
                                                                       uint32_t M, uint32_t N, float* grad_sigmas, float* grad_rgbs) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = rays[n * 3], offset = rays[n * 3 + 1], num_steps = rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps >= M) return;
    const float g0 = grad_img[index * 3], g1 = grad_img[index * 3 + 1], g2 = grad_img[index * 3 + 2], gw = grad_ws[index];
    const float rf = image[index * 3], gf = image[index * 3 + 1], bf = image[index * 3 + 2], Tf = 1 - weights_sum[index];
    float T = 1.0f, r = 0, g = 0, b = 0;
    for (uint32_t s = 0; s < num_steps; ++s) {                            // no early out here (:361)
        const size_t i = offset + s;
        const float alpha = sigmas[i], w = alpha * T;
        const float c0 = rgbs[3 * i], c1 = rgbs[3 * i + 1], c2 = rgbs[3 * i + 2];
        r = fmaf(w, c0, r); g = fmaf(w, c1, g); b = fmaf(w, c2, b);
        T *= 1.0f - alpha;                                               // T(t+1)
        grad_rgbs[3 * i] = g0 * w; grad_rgbs[3 * i + 1] = g1 * w; grad_rgbs[3 * i + 2] = g2 * w;
        float acc = g0 * fmaf(T, c0, -(rf - r));
        acc = fmaf(g1, fmaf(T, c1, -(gf - g)), acc);
        acc = fmaf(g2, fmaf(T, c2, -(bf - b)), acc);
        acc = fmaf(gw, Tf, acc);
        grad_sigmas[i] = deltas[i] * acc;
    }
}

__global__ void __launch_bounds__(256) march_rays_kernel(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
                                                         const float* __restrict__ rays_t, const float* __restrict__ rays_o,
                                                         const float* __restrict__ rays_d, float bound, uint32_t H,
                                                         const float* __restrict__ grid, float mean_density, const float* __restrict__ nears,
                                                         const float* __restrict__ fars, float* xyzs, float* dirs, float* deltas,
                                                         uint32_t perturb) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    float t = rays_t[n];
    const GridStepper g(rays_o + 3 * (size_t)index, rays_d + 3 * (size_t)index, bound, H, grid, mean_density);
    const float far = fars[index];
    (void)nears;
    float* px = xyzs + (size_t)n * n_step * 3; float* pd = dirs + (size_t)n * n_step * 3; float* pt = deltas + (size_t)n * n_step * 2;
    if (perturb) { Pcg32 rng((uint64_t)n, (uint64_t)perturb); t = fmaf(g.dt_min, rng.next_float(), t); }
    float last_t = t, x, y, z, t_exit;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (g.probe(t, x, y, z, t_exit)) {
            px[0] = x; px[1] = y; px[2] = z; pd[0] = g.dx; pd[1] = g.dy; pd[2] = g.dz;
            const float dt = g.step_size(t);
            t += dt;
            pt[0] = dt; pt[1] = t - last_t;                                 // second delta: real advance, for depth
            last_t = t;
            px += 3; pd += 3; pt += 2; ++step;
        } else t = g.skip(t, t_exit);
    }
}

__global__ void __launch_bounds__(256) composite_rays_kernel(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive, float* rays_t,
                                                             const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                             const float* __restrict__ normals, const float* __restrict__ deltas,
                                                             float* weights_sum, float* depth, float* image, float* normal_map) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    float t = rays_t[n];
    const size_t base = (size_t)n * n_step;
    float ws = weights_sum[index], d = depth[index];
    float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
    float nx = normal_map[index * 3], ny = normal_map[index * 3 + 1], nz = normal_map[index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        const size_t i = base + step;
        if (deltas[2 * i] == 0) break;                                      // unused slot: the ray left the volume
        const float alpha = sigmas[i], T = 1 - ws, w = alpha * T;
        ws += w;
        t += deltas[2 * i + 1];
        d = fmaf(w, t, d);
        r = fmaf(w, rgbs[3 * i], r); g = fmaf(w, rgbs[3 * i + 1], g); b = fmaf(w, rgbs[3 * i + 2], b);
        nx = fmaf(w, normals[3 * i], nx); ny = fmaf(w, normals[3 * i + 1], ny); nz = fmaf(w, normals[3 * i + 2], nz);
        if ((double)T < 1e-2) break;                                        // ray is opaque: retire it (:676; double literal)
        ++step;
    }
    rays_t[n] = step < n_step ? -1.0f : t;
    weights_sum[index] = ws; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    normal_map[index * 3] = nx; normal_map[index * 3 + 1] = ny; normal_map[index * 3 + 2] = nz;
}

__global__ void __launch_bounds__(256) compact_rays_kernel(uint32_t n_alive, int* rays_alive, const int* __restrict__ rays_alive_old, float* rays_t,
                                                           const float* __restrict__ rays_t_old, int* alive_counter) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    if (rays_t_old[n] >= 0) {                                               // still alive (:741)
        const int slot = atomicAdd(alive_counter, 1);
        rays_alive[slot] = rays_alive_old[n];
        rays_t[slot] = rays_t_old[n];
    }
}

inline uint32_t blocks(uint32_t n) { return (n + 255) / 256; }

}  // namespace

extern "C" {

int ac_march_rays_train(const float* rays_o, const float* rays_d, const float* grid, float mean_density, int iter_density, float bound,
                        uint32_t N, uint32_t H, uint32_t M, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                        uint32_t perturb, void* stream) {
    (void)iter_density;
    if (!rays_o || !rays_d || !grid || !xyzs || !dirs || !deltas || !rays || !counter || H < 2) return AC_E_INVALID_ARG;
    if (N == 0) return AC_OK;
    march_rays_train_kernel<<<blocks(N), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, grid, mean_density, bound, N, H, M, xyzs, dirs, deltas, rays,
                                                                       counter, perturb);
    return acb::launched();
}

int ac_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, float bound, uint32_t M,
                                    uint32_t N, float* weights_sum, float* image, void* stream) {
    (void)deltas; (void)bound;
    if (!sigmas || !rgbs || !rays || !weights_sum || !image) return AC_E_INVALID_ARG;
    if (N == 0) return AC_OK;
    composite_train_forward_kernel<<<blocks(N), 256, 0, (cudaStream_t)stream>>>(sigmas, rgbs, rays, M, N, weights_sum, image);
    return acb::launched();
}

int ac_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas, const float* rgbs,
                                     const float* deltas, const int32_t* rays, const float* weights_sum, const float* image, float bound,
                                     uint32_t M, uint32_t N, float* grad_sigmas, float* grad_rgbs, void* stream) {
    (void)bound;
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays || !weights_sum || !image || !grad_sigmas || !grad_rgbs)
        return AC_E_INVALID_ARG;
    if (N == 0) return AC_OK;
    composite_train_backward_kernel<<<blocks(N), 256, 0, (cudaStream_t)stream>>>(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum,
                                                                               image, M, N, grad_sigmas, grad_rgbs);
    return acb::launched();
}

int ac_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                  float bound, uint32_t H, const float* grid, float mean_density, const float* nears, const float* fars, float* xyzs,
                  float* dirs, float* deltas, uint32_t perturb, void* stream) {
    if (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !nears || !fars || !xyzs || !dirs || !deltas || H < 2) return AC_E_INVALID_ARG;
    if (n_alive == 0) return AC_OK;
    march_rays_kernel<<<blocks(n_alive), 256, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, H, grid,
                                                                       mean_density, nears, fars, xyzs, dirs, deltas, perturb);
    return acb::launched();
}

int ac_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs,
                      const float* normals, const float* deltas, float* weights_sum, float* depth, float* image, float* normal_map,
                      void* stream) {
    if (!rays_alive || !rays_t || !sigmas || !rgbs || !normals || !deltas || !weights_sum || !depth || !image || !normal_map)
        return AC_E_INVALID_ARG;
    if (n_alive == 0) return AC_OK;
    composite_rays_kernel<<<blocks(n_alive), 256, 0, (cudaStream_t)stream>>>(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, normals, deltas,
                                                                           weights_sum, depth, image, normal_map);
    return acb::launched();
}

int ac_compact_rays(uint32_t n_alive, int32_t* rays_alive, const int32_t* rays_alive_old, float* rays_t, const float* rays_t_old,
                    int32_t* alive_counter, void* stream) {
    if (!rays_alive || !rays_alive_old || !rays_t || !rays_t_old || !alive_counter) return AC_E_INVALID_ARG;
    if (n_alive == 0) return AC_OK;
    compact_rays_kernel<<<blocks(n_alive), 256, 0, (cudaStream_t)stream>>>(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter);
    return acb::launched();
}

}  // extern "C"
