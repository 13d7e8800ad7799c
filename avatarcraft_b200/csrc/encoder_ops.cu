// encoder_ops.cu -- drop-in replacements for the reference's encoder operators:
//   hash_encode_forward / hash_encode_backward (encoder/hashencoder/src/hashencoder.cu)
//   sh_encode_forward / sh_encode_backward     (encoder/shencoder/src/shencoder.cu)
// Same buffer layouts and argument meaning; fp32; launched on the caller's stream.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"
#include "nsr_device.cuh"

using namespace acb;

namespace {

template <uint32_t D>
__device__ __forceinline__ uint32_t corner_slot(const LevelMeta& m, const uint32_t (&cell)[D]) {
    // dense walk while the running stride fits, else xor-hash (hashencoder.cu:54-70, :35-51)
    constexpr uint32_t mult[3] = {1u, 2654435761u, 805459861u};
    uint32_t stride = 1, slot = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (stride <= m.size) { slot += cell[d] * stride; stride *= m.res1; }
    }
    if (stride > m.size) {
        slot = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) slot ^= cell[d] * mult[d];
    }
    return slot % m.size;
}

// One thread per (point, level); level = blockIdx.y so one level's table slice is hot in
// L2/L1 at a time.  Output layout [L,B,C] as the reference wrapper expects (hashgrid.py:31).
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) hash_forward_kernel(const float* __restrict__ inputs, const float* __restrict__ table,
                                                           const int32_t* __restrict__ offsets, float* __restrict__ outputs,
                                                           uint32_t B, uint32_t L, float S, uint32_t H, bool want_jac,
                                                           float* __restrict__ dy_dx, int32_t* __restrict__ corner_ids, bool point_major) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelMeta m = make_level_meta(offsets, level, S, H, D);
    float x[D];
    bool inside = true;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) { x[d] = inputs[(size_t)b * D + d]; inside = inside && !(x[d] < 0.f || x[d] > 1.f); }
    // reference layout [L,B,C] (hashgrid.py:31) or point-major [B,L,C] = the [B, L*C] tensor the module returns
    float* out = outputs + (point_major ? ((size_t)b * L + level) : ((size_t)level * B + b)) * C;
    float* jac = want_jac ? dy_dx + (((size_t)b * L + level) * D) * C : nullptr;
    int32_t* ids = corner_ids ? corner_ids + ((size_t)level * B + b) * (1u << D) : nullptr;
    if (!inside) {
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) out[c] = 0.f;
        if (jac) for (uint32_t i = 0; i < D * C; ++i) jac[i] = 0.f;
        if (ids) for (uint32_t k = 0; k < (1u << D); ++k) ids[k] = -1;
        return;
    }
    const float* __restrict__ tab = table + (size_t)m.offset * C;
    float frac[D];
    uint32_t base[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float p = fmaf(x[d], m.scale, 0.5f);
        const float fl = floorf(p);
        base[d] = (uint32_t)fl;
        frac[d] = p - fl;
    }
    float acc[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) acc[c] = 0.f;
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) {
        float w = 1.f;
        uint32_t cell[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            if (k & (1u << d)) { w *= frac[d]; cell[d] = base[d] + 1u; }
            else { w *= 1.f - frac[d]; cell[d] = base[d]; }
        }
        const uint32_t slot = corner_slot<D>(m, cell);
        if (ids) ids[k] = (int32_t)slot;
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) acc[c] = fmaf(w, __ldg(tab + (size_t)slot * C + c), acc[c]);
    }
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
    if (jac) {   // d(out)/d(x_g): difference across axis g, blended over the other axes (hashencoder.cu:176-218)
#pragma unroll
        for (uint32_t g = 0; g < D; ++g) {
            float dacc[C];
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) dacc[c] = 0.f;
#pragma unroll
            for (uint32_t k = 0; k < (1u << (D - 1)); ++k) {
                float w = m.scale;
                uint32_t cell[D];
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; ++nd) {
                    const uint32_t d = nd >= g ? nd + 1 : nd;
                    if (k & (1u << nd)) { w *= frac[d]; cell[d] = base[d] + 1u; }
                    else { w *= 1.f - frac[d]; cell[d] = base[d]; }
                }
                cell[g] = base[g];
                const uint32_t lo = corner_slot<D>(m, cell);
                cell[g] = base[g] + 1u;
                const uint32_t hi = corner_slot<D>(m, cell);
#pragma unroll
                for (uint32_t c = 0; c < C; ++c)
                    dacc[c] = fmaf(w, __ldg(tab + (size_t)hi * C + c) - __ldg(tab + (size_t)lo * C + c), dacc[c]);
            }
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) jac[g * C + c] = dacc[c];
        }
    }
}

// Scatter w * grad into the 2^D corners (hashencoder.cu:223-308).  fp32 reductions
// (red.global.add.f32; vectorised .v2 when C is even) -- no return value needed.
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) hash_backward_kernel(const float* __restrict__ grad, const float* __restrict__ inputs,
                                                            const int32_t* __restrict__ offsets, float* __restrict__ grad_table,
                                                            uint32_t B, uint32_t L, float S, uint32_t H, bool point_major) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const LevelMeta m = make_level_meta(offsets, level, S, H, D);
    float frac[D];
    uint32_t base[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        const float x = inputs[(size_t)b * D + d];
        if (x < 0.f || x > 1.f) return;
        const float p = fmaf(x, m.scale, 0.5f);
        const float fl = floorf(p);
        base[d] = (uint32_t)fl;
        frac[d] = p - fl;
    }
    float g[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) g[c] = grad[(point_major ? ((size_t)b * L + level) : ((size_t)level * B + b)) * C + c];
    float* __restrict__ dst = grad_table + (size_t)m.offset * C;
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) {
        float w = 1.f;
        uint32_t cell[D];
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            if (k & (1u << d)) { w *= frac[d]; cell[d] = base[d] + 1u; }
            else { w *= 1.f - frac[d]; cell[d] = base[d]; }
        }
        const uint32_t slot = corner_slot<D>(m, cell);
        if (C % 2 == 0) {
#pragma unroll
            for (uint32_t c = 0; c < C; c += 2)
                atomicAdd(reinterpret_cast<float2*>(dst + (size_t)slot * C + c), make_float2(w * g[c], w * g[c + 1]));
        } else {
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) atomicAdd(dst + (size_t)slot * C + c, w * g[c]);
        }
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (hashencoder.cu:311-337)
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) hash_input_backward_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                                                  float* __restrict__ grad_inputs, uint32_t B, uint32_t L, bool point_major) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* jac = dy_dx + (size_t)b * L * D * C;
    float r = 0.f;
    for (uint32_t l = 0; l < L; ++l)
#pragma unroll
        for (uint32_t c = 0; c < C; ++c)
            r = fmaf(grad[(point_major ? ((size_t)b * L + l) : ((size_t)l * B + b)) * C + c], jac[((size_t)l * D + d) * C + c], r);
    grad_inputs[t] = r;
}

__global__ void level_scales_kernel(float* scales, uint32_t L, float S, uint32_t H) {
    const uint32_t l = threadIdx.x;
    if (l < L) scales[l] = fmaf(exp2f((float)l * S), (float)H, -1.0f);
}

template <uint32_t D>
int launch_forward(uint32_t C, dim3 grid, cudaStream_t st, const float* in, const float* tab, const int32_t* off, float* out,
                   uint32_t B, uint32_t L, float S, uint32_t H, bool jac, float* dy_dx, int32_t* ids, bool pm) {
    switch (C) {
        case 1: hash_forward_kernel<D, 1><<<grid, 256, 0, st>>>(in, tab, off, out, B, L, S, H, jac, dy_dx, ids, pm); break;
        case 2: hash_forward_kernel<D, 2><<<grid, 256, 0, st>>>(in, tab, off, out, B, L, S, H, jac, dy_dx, ids, pm); break;
        case 4: hash_forward_kernel<D, 4><<<grid, 256, 0, st>>>(in, tab, off, out, B, L, S, H, jac, dy_dx, ids, pm); break;
        case 8: hash_forward_kernel<D, 8><<<grid, 256, 0, st>>>(in, tab, off, out, B, L, S, H, jac, dy_dx, ids, pm); break;
        default: return AC_E_UNSUPPORTED;
    }
    return acb::launched();
}

template <uint32_t D>
int launch_backward(uint32_t C, dim3 grid, cudaStream_t st, const float* grad, const float* in, const int32_t* off, float* gt,
                    uint32_t B, uint32_t L, float S, uint32_t H, bool jac, const float* dy_dx, float* gi, bool pm) {
    const uint32_t gin = (B * D + 255) / 256;
    switch (C) {
        case 1: hash_backward_kernel<D, 1><<<grid, 256, 0, st>>>(grad, in, off, gt, B, L, S, H, pm);
                if (jac) hash_input_backward_kernel<D, 1><<<gin, 256, 0, st>>>(grad, dy_dx, gi, B, L, pm); break;
        case 2: hash_backward_kernel<D, 2><<<grid, 256, 0, st>>>(grad, in, off, gt, B, L, S, H, pm);
                if (jac) hash_input_backward_kernel<D, 2><<<gin, 256, 0, st>>>(grad, dy_dx, gi, B, L, pm); break;
        case 4: hash_backward_kernel<D, 4><<<grid, 256, 0, st>>>(grad, in, off, gt, B, L, S, H, pm);
                if (jac) hash_input_backward_kernel<D, 4><<<gin, 256, 0, st>>>(grad, dy_dx, gi, B, L, pm); break;
        case 8: hash_backward_kernel<D, 8><<<grid, 256, 0, st>>>(grad, in, off, gt, B, L, S, H, pm);
                if (jac) hash_input_backward_kernel<D, 8><<<gin, 256, 0, st>>>(grad, dy_dx, gi, B, L, pm); break;
        default: return AC_E_UNSUPPORTED;
    }
    int rc = acb::launched();
    if (rc == AC_OK && jac) rc = acb::launched();
    return rc;
}

}  // namespace

extern "C" {

static int hash_forward_any(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                            uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float* dy_dx,
                            int32_t* corner_ids, void* stream, bool pm) {
    if (!inputs || !embeddings || !offsets || !outputs || (calc_grad_inputs && !dy_dx) || L == 0) return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    const dim3 grid((B + 255) / 256, L, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) return launch_forward<2>(C, grid, st, inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx, corner_ids, pm);
    if (D == 3) return launch_forward<3>(C, grid, st, inputs, embeddings, offsets, outputs, B, L, S, H, calc_grad_inputs != 0, dy_dx, corner_ids, pm);
    return AC_E_UNSUPPORTED;
}
int ac_hash_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                           uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float* dy_dx,
                           int32_t* corner_ids, void* stream) {
    return hash_forward_any(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, corner_ids, stream, false);
}
int ac_hash_encode_forward_pm(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                              uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, float* dy_dx, void* stream) {
    return hash_forward_any(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx, nullptr, stream, true);
}

static int hash_backward_any(const float* grad, const float* inputs, const int32_t* offsets,
                             float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             int calc_grad_inputs, const float* dy_dx, float* grad_inputs, void* stream, bool pm) {
    if (!grad || !inputs || !offsets || !grad_embeddings || (calc_grad_inputs && (!dy_dx || !grad_inputs)) || L == 0)
        return AC_E_INVALID_ARG;
    if (B == 0) return AC_OK;
    const dim3 grid((B + 255) / 256, L, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) return launch_backward<2>(C, grid, st, grad, inputs, offsets, grad_embeddings, B, L, S, H, calc_grad_inputs != 0, dy_dx, grad_inputs, pm);
    if (D == 3) return launch_backward<3>(C, grid, st, grad, inputs, offsets, grad_embeddings, B, L, S, H, calc_grad_inputs != 0, dy_dx, grad_inputs, pm);
    return AC_E_UNSUPPORTED;
}
int ac_hash_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                            float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            int calc_grad_inputs, const float* dy_dx, float* grad_inputs, void* stream) {
    (void)embeddings;
    return hash_backward_any(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_inputs, stream, false);
}
int ac_hash_encode_backward_pm(const float* grad, const float* inputs, const int32_t* offsets, float* grad_embeddings, uint32_t B,
                               uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs, const float* dy_dx,
                               float* grad_inputs, void* stream) {
    return hash_backward_any(grad, inputs, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs, dy_dx, grad_inputs, stream, true);
}

int ac_hash_level_scales(float* scales, uint32_t L, float S, uint32_t H, void* stream) {
    if (!scales || L == 0 || L > 32) return AC_E_INVALID_ARG;
    level_scales_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scales, L, S, H);
    return acb::launched();
}

}  // extern "C"
