// sh_ops.cu -- real spherical-harmonics direction encoder, degree <= 8 (64 coefficients) and its input
// gradient.  Replaces sh_encode_forward / sh_encode_backward (encoder/shencoder/src/shencoder.cu:28-400).
//
// The reference hard-codes 64 sympy-generated polynomials; here the same polynomials (for arbitrary, not
// necessarily unit, xyz) come from the factorisation
//     Y_l^{+m} = K_l^m * Q_l^m(z) * Re (x+iy)^m,   Y_l^{-m} = K_l^m * Q_l^m(z) * Im (x+iy)^m,   Y_l^0 = N_l^0 Q_l^0(z)
// with Q_l^m = d^m P_l / dz^m built by the Legendre three-term recurrence and
// K_l^m = (-1)^m sqrt(2) sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) (Condon-Shortley phase, as the reference's signs).
// Output index l*l + l + m; dy_dx [B, 3, degree^2] (d/dx block, d/dy block, d/dz block) as the reference.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/avatarcraft_b200.h"
#include "launch_util.cuh"

namespace {

struct ShConst { float k[8][8]; };     // k[l][m], m <= l

ShConst make_consts() {
    ShConst c{};
    for (int l = 0; l < 8; ++l)
        for (int m = 0; m <= l; ++m) {
            double f = 1.0;
            for (int i = l - m + 1; i <= l + m; ++i) f *= (double)i;           // (l+m)!/(l-m)!
            double n = sqrt((2.0 * l + 1.0) / (4.0 * M_PI) / f);
            c.k[l][m] = (float)(m == 0 ? n : ((m & 1) ? -1.0 : 1.0) * sqrt(2.0) * n);
        }
    return c;
}

__global__ void __launch_bounds__(256) sh_forward_kernel(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B,
                                                         uint32_t degree, bool want_jac, float* __restrict__ dy_dx, const ShConst K) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float x = inputs[3 * (size_t)b], y = inputs[3 * (size_t)b + 1], z = inputs[3 * (size_t)b + 2];
    const uint32_t n = degree * degree;
    float* out = outputs + (size_t)b * n;
    float* dx = want_jac ? dy_dx + (size_t)b * 3 * n : nullptr;
    float* dy = dx ? dx + n : nullptr;
    float* dz = dx ? dy + n : nullptr;
    float cm1 = 0.f, sm1 = 0.f, cm = 1.f, sm = 0.f;       // Re/Im (x+iy)^(m-1), (x+iy)^m
    float dfact = 1.f;                                    // (2m-1)!!
    for (uint32_t m = 0; m < degree; ++m) {
        if (m > 0) {
            const float c2 = x * cm - y * sm, s2 = x * sm + y * cm;
            cm1 = cm; sm1 = sm; cm = c2; sm = s2;
            dfact *= (float)(2 * m - 1);
        }
        // Q_l^m and Q_l^{m+1} for l = m .. degree-1, marching both recurrences together
        float q_prev2 = 0.f, q_prev = 0.f;                // Q_{l-2}^m, Q_{l-1}^m
        float r_prev2 = 0.f, r_prev = 0.f;                // Q_{l-2}^{m+1}, Q_{l-1}^{m+1}
        const float dfact1 = dfact * (float)(2 * m + 1);  // (2m+1)!!
        for (uint32_t l = m; l < degree; ++l) {
            float q, r;
            if (l == m) q = dfact;
            else if (l == m + 1) q = (float)(2 * m + 1) * z * q_prev;
            else q = ((float)(2 * l - 1) * z * q_prev - (float)(l + m - 1) * q_prev2) / (float)(l - m);
            if (l == m) r = 0.f;
            else if (l == m + 1) r = dfact1;
            else if (l == m + 2) r = (float)(2 * m + 3) * z * r_prev;
            else r = ((float)(2 * l - 1) * z * r_prev - (float)(l + m) * r_prev2) / (float)(l - m - 1);
            const float k = K.k[l][m];
            const uint32_t base = l * l + l;
            if (m == 0) {
                out[base] = k * q;
                if (dx) { dx[base] = 0.f; dy[base] = 0.f; dz[base] = k * r; }
            } else {
                const float kq = k * q, fm = (float)m;
                out[base + m] = kq * cm;
                out[base - m] = kq * sm;
                if (dx) {
                    dx[base + m] = kq * fm * cm1;  dy[base + m] = -kq * fm * sm1;  dz[base + m] = k * r * cm;
                    dx[base - m] = kq * fm * sm1;  dy[base - m] = kq * fm * cm1;   dz[base - m] = k * r * sm;
                }
            }
            q_prev2 = q_prev; q_prev = q; r_prev2 = r_prev; r_prev = r;
        }
    }
}

// grad_inputs[b,d] = sum_c grad[b,c] * dy_dx[b,d,c]   (shencoder.cu:360-384; the reference accumulates with +=
// into a zero-initialised buffer, so plain assignment is equivalent)
__global__ void __launch_bounds__(256) sh_backward_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                                          float* __restrict__ grad_inputs, uint32_t B, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const uint32_t b = t / 3, d = t - 3 * b;
    const float* g = grad + (size_t)b * n;
    const float* j = dy_dx + ((size_t)b * 3 + d) * n;
    float r = 0.f;
    for (uint32_t c = 0; c < n; ++c) r = fmaf(g[c], j[c], r);
    grad_inputs[t] += r;
}

}  // namespace

extern "C" {

int ac_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, int calc_grad_inputs, float* dy_dx,
                         void* stream) {
    if (!inputs || !outputs || (calc_grad_inputs && !dy_dx)) return AC_E_INVALID_ARG;
    if (D != 3 || degree < 1 || degree > 8) return AC_E_UNSUPPORTED;
    if (B == 0) return AC_OK;
    static const ShConst K = make_consts();
    sh_forward_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(inputs, outputs, B, degree, calc_grad_inputs != 0, dy_dx, K);
    return acb::launched();
}

int ac_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                          float* grad_inputs, void* stream) {
    (void)inputs;
    if (!grad || !dy_dx || !grad_inputs) return AC_E_INVALID_ARG;
    if (D != 3 || degree < 1 || degree > 8) return AC_E_UNSUPPORTED;
    if (B == 0) return AC_OK;
    sh_backward_kernel<<<(B * 3 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad, dy_dx, grad_inputs, B, degree * degree);
    return acb::launched();
}

}  // extern "C"
