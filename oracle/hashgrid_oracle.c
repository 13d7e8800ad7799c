/*
 * oracle/hashgrid_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, OpenMP) of the reference's multi-resolution hash-grid
 * encoder.  Nothing in the product path (avatarcraft_b200/) may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg use it, and only as the checker / CPU baseline.
 *
 * Reference being restated (paths relative to /root/reference):
 *   encoder/hashencoder/src/hashencoder.cu:35-51   spatial hash (xor of coord*prime, primes[0]==1)
 *   encoder/hashencoder/src/hashencoder.cu:54-70   dense-vs-hashed corner index, final "% hashmap_size"
 *   encoder/hashencoder/src/hashencoder.cu:94-119  inputs outside [0,1] -> zero features (and zero dy_dx)
 *   encoder/hashencoder/src/hashencoder.cu:120-172 level scale / resolution, cell position, 2^D-corner blend
 *   encoder/hashencoder/src/hashencoder.cu:176-218 dy_dx (derivative of the blend wrt each input coordinate)
 *   encoder/hashencoder/src/hashencoder.cu:223-308 backward scatter into the table
 *   encoder/hashencoder/src/hashencoder.cu:311-337 input gradient from dy_dx
 *
 * Parity pin: the reference ships no golden vectors for this path.  The
 * restatement is pinned (a) on the GPU box against the reference's own hashencoder.cu
 * compiled into oracle/_ref/ (tests/test_gpu_reference_kernel.py) and (b) through the
 * imported reference model in oracle/make_golden.py.
 *
 * Floating-point notes.  nvcc (default -fmad=true) contracts "x*scale+0.5f" in the
 * reference kernel into one FMA, so the cell position here uses fmaf().  exp2f on the GPU
 * is MUFU-based and may differ from glibc's by an ulp; callers can therefore pass the
 * per-level scales explicitly (scales != NULL), e.g. read back from the device.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_D 3
#define ORACLE_MAX_C 8

static uint32_t corner_slot(uint32_t D, uint32_t level_entries, uint32_t res, const uint32_t *cell)
{
    static const uint32_t mult[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                     2097192037u, 1434869437u, 2165219737u};
    uint32_t stride = 1, slot = 0, d = 0;
    while (d < D && stride <= level_entries) {   /* dense walk while the running stride still fits */
        slot += cell[d] * stride;
        stride *= (res + 1u);
        ++d;
    }
    if (stride > level_entries) {                /* too fine for a dense grid: spatial hash */
        slot = 0;
        for (d = 0; d < D; ++d) slot ^= cell[d] * mult[d];
    }
    return slot % level_entries;
}

float oracle_level_scale(uint32_t level, float S, uint32_t H)
{
    return exp2f((float)level * S) * (float)H - 1.0f;
}

/* inputs [B,D] in [0,1]; table [n_entries,C]; offsets [L+1]; outputs [L,B,C];
 * dy_dx [B,L,D,C] (only if calc_grad_inputs); corner_ids [L,B,2^D] optional (entry index
 * inside the level, before the *C channel stride); scales [L] optional override. */
void oracle_hashgrid_forward(const float *inputs, const float *table, const int32_t *offsets,
                             float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                             float S, uint32_t H, int calc_grad_inputs, float *dy_dx,
                             int32_t *corner_ids, const float *scales)
{
    const uint32_t ncorner = 1u << D;
#pragma omp parallel for schedule(static)
    for (int64_t bi = 0; bi < (int64_t)B; ++bi) {
        const uint32_t b = (uint32_t)bi;
        const float *x = inputs + (size_t)b * D;
        int inside = 1;
        for (uint32_t d = 0; d < D; ++d)
            if (x[d] < 0.0f || x[d] > 1.0f) inside = 0;
        for (uint32_t l = 0; l < L; ++l) {
            float *out = outputs + ((size_t)l * B + b) * C;
            float *jac = calc_grad_inputs ? dy_dx + (((size_t)b * L + l) * D) * C : NULL;
            int32_t *ids = corner_ids ? corner_ids + ((size_t)l * B + b) * ncorner : NULL;
            if (!inside) {
                for (uint32_t c = 0; c < C; ++c) out[c] = 0.0f;
                if (jac) memset(jac, 0, sizeof(float) * D * C);
                if (ids) for (uint32_t k = 0; k < ncorner; ++k) ids[k] = -1;
                continue;
            }
            const float *tab = table + (size_t)(uint32_t)offsets[l] * C;
            const uint32_t entries = (uint32_t)(offsets[l + 1] - offsets[l]);
            const float scale = scales ? scales[l] : oracle_level_scale(l, S, H);
            const uint32_t res = (uint32_t)ceilf(scale) + 1u;
            float frac[ORACLE_MAX_D];
            uint32_t base[ORACLE_MAX_D];
            for (uint32_t d = 0; d < D; ++d) {
                float p = fmaf(x[d], scale, 0.5f);
                float fl = floorf(p);
                base[d] = (uint32_t)fl;
                frac[d] = p - (float)base[d];
            }
            float acc[ORACLE_MAX_C] = {0};
            for (uint32_t k = 0; k < ncorner; ++k) {
                float w = 1.0f;
                uint32_t cell[ORACLE_MAX_D];
                for (uint32_t d = 0; d < D; ++d) {
                    if (k & (1u << d)) { w *= frac[d];        cell[d] = base[d] + 1u; }
                    else               { w *= 1.0f - frac[d]; cell[d] = base[d]; }
                }
                uint32_t slot = corner_slot(D, entries, res, cell);
                if (ids) ids[k] = (int32_t)slot;
                for (uint32_t c = 0; c < C; ++c) acc[c] = fmaf(w, tab[(size_t)slot * C + c], acc[c]);
            }
            for (uint32_t c = 0; c < C; ++c) out[c] = acc[c];
            if (jac) {
                for (uint32_t g = 0; g < D; ++g) {
                    float dacc[ORACLE_MAX_C] = {0};
                    for (uint32_t k = 0; k < (1u << (D - 1)); ++k) {
                        float w = scale;
                        uint32_t cell[ORACLE_MAX_D];
                        for (uint32_t nd = 0; nd < D - 1; ++nd) {
                            uint32_t d = nd >= g ? nd + 1 : nd;
                            if (k & (1u << nd)) { w *= frac[d];        cell[d] = base[d] + 1u; }
                            else                { w *= 1.0f - frac[d]; cell[d] = base[d]; }
                        }
                        cell[g] = base[g];
                        uint32_t lo = corner_slot(D, entries, res, cell);
                        cell[g] = base[g] + 1u;
                        uint32_t hi = corner_slot(D, entries, res, cell);
                        for (uint32_t c = 0; c < C; ++c)
                            dacc[c] += w * (tab[(size_t)hi * C + c] - tab[(size_t)lo * C + c]);
                    }
                    for (uint32_t c = 0; c < C; ++c) jac[g * C + c] = dacc[c];
                }
            }
        }
    }
}

/* grad [L,B,C]; grad_table [n_entries,C] is ACCUMULATED into (caller zeroes it, as the
 * reference's zeros_like does, hashgrid.py:59).  Accumulation is sequential in b and done
 * in double, so it is deterministic and at least as accurate as the reference's fp32
 * atomics (whose order is unspecified). */
void oracle_hashgrid_backward(const float *grad, const float *inputs, const int32_t *offsets,
                              float *grad_table, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                              float S, uint32_t H, int calc_grad_inputs, const float *dy_dx,
                              float *grad_inputs, const float *scales)
{
    const uint32_t ncorner = 1u << D;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t li = 0; li < (int64_t)L; ++li) {
        const uint32_t l = (uint32_t)li;
        const uint32_t entries = (uint32_t)(offsets[l + 1] - offsets[l]);
        double *accum = (double *)calloc((size_t)entries * C, sizeof(double));
        const float scale = scales ? scales[l] : oracle_level_scale(l, S, H);
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        for (uint32_t b = 0; b < B; ++b) {
            const float *x = inputs + (size_t)b * D;
            int inside = 1;
            for (uint32_t d = 0; d < D; ++d)
                if (x[d] < 0.0f || x[d] > 1.0f) inside = 0;
            if (!inside) continue;
            float frac[ORACLE_MAX_D];
            uint32_t base[ORACLE_MAX_D];
            for (uint32_t d = 0; d < D; ++d) {
                float p = fmaf(x[d], scale, 0.5f);
                float fl = floorf(p);
                base[d] = (uint32_t)fl;
                frac[d] = p - (float)base[d];
            }
            const float *g = grad + ((size_t)l * B + b) * C;
            for (uint32_t k = 0; k < ncorner; ++k) {
                float w = 1.0f;
                uint32_t cell[ORACLE_MAX_D];
                for (uint32_t d = 0; d < D; ++d) {
                    if (k & (1u << d)) { w *= frac[d];        cell[d] = base[d] + 1u; }
                    else               { w *= 1.0f - frac[d]; cell[d] = base[d]; }
                }
                uint32_t slot = corner_slot(D, entries, res, cell);
                for (uint32_t c = 0; c < C; ++c)
                    accum[(size_t)slot * C + c] += (double)(w * g[c]);
            }
        }
        float *dst = grad_table + (size_t)(uint32_t)offsets[l] * C;
        for (size_t i = 0; i < (size_t)entries * C; ++i) dst[i] += (float)accum[i];
        free(accum);
    }
    if (calc_grad_inputs) {
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < (int64_t)B * D; ++t) {
            uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t % D);
            const float *jac = dy_dx + (size_t)b * L * D * C;
            float r = 0.0f;
            for (uint32_t l = 0; l < L; ++l)
                for (uint32_t c = 0; c < C; ++c)
                    r += grad[((size_t)l * B + b) * C + c] * jac[((size_t)l * D + d) * C + c];
            grad_inputs[t] = r;
        }
    }
}
