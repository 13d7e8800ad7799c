/*
 * avatarcraft_b200.h -- C ABI of libavatarcraft_b200.so (sm_100a).
 *
 * The drop-in boundary for the AvatarCraft hot path: every entry point takes plain
 * device pointers, sizes and a CUDA stream (passed as void* = cudaStream_t); no torch or
 * C++ types cross it.  The caller owns and allocates every buffer (as the reference's
 * Python wrappers do, encoder/hashencoder/hashgrid.py:31-38,59-66); the library never
 * allocates device memory.  Every function returns 0 on success or a negative AC_E_* code
 * and never synchronises the stream.  All kernels are launched on `stream` and are
 * CUDA-graph capturable.  Citations are file:line inside the reference repository
 * (songrise/AvatarCraft @ 190a4bb).
 */
#ifndef AVATARCRAFT_B200_H
#define AVATARCRAFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AC_OK 0
#define AC_E_INVALID_ARG (-1)    /* NULL pointer, bad size (reference: TORCH_CHECK -> RuntimeError) */
#define AC_E_UNSUPPORTED (-2)    /* D not in {2,3} / C not in {1,2,4,8} (hashencoder.cu:349,364)      */
#define AC_E_CUDA (-3)           /* a CUDA launch failed; see ac_last_cuda_error()                     */
#define AC_E_WORKSPACE (-4)      /* workspace too small                                                */

/* Library identity: version string and the CUDA arch it was compiled for ("sm_100a"). */
const char *ac_version(void);
const char *ac_last_cuda_error(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t ac_launch_count(void);
/* A caller that replays a captured CUDA graph of this library's launches adds the graph's kernel count per replay. */
void ac_launch_count_add(uint64_t n);

/* --------------------------------------------------------------------------------------
 * Multi-resolution hash-grid encoder.  Replaces the pybind ops
 *   hash_encode_forward  (encoder/hashencoder/src/hashencoder.h:13, hashencoder.cu:413-436)
 *   hash_encode_backward (encoder/hashencoder/src/hashencoder.h:14, hashencoder.cu:439-470)
 * with the same argument order and buffer layouts:
 *   inputs [B,D] in [0,1]; embeddings [n_entries,C]; offsets [L+1] int32;
 *   outputs [L,B,C]; dy_dx [B,L,D,C] (written only when calc_grad_inputs != 0);
 *   grad [L,B,C]; grad_embeddings [n_entries,C] (accumulated into; caller zeroes);
 *   grad_inputs [B,D].  S = log2(per_level_scale), H = base resolution.  fp32 only.
 * corner_ids (optional, may be NULL) receives the table slot of every corner,
 * [L,B,2^D] int32 (-1 for out-of-range inputs) -- the integer path parity tests pin.
 * ------------------------------------------------------------------------------------ */
int ac_hash_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                           float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                           float S, uint32_t H, int calc_grad_inputs, float *dy_dx,
                           int32_t *corner_ids, void *stream);
int ac_hash_encode_backward(const float *grad, const float *inputs, const float *embeddings,
                            const int32_t *offsets, float *grad_embeddings, uint32_t B, uint32_t D,
                            uint32_t C, uint32_t L, float S, uint32_t H, int calc_grad_inputs,
                            const float *dy_dx, float *grad_inputs, void *stream);
/* Point-major variants: outputs / grad are [B, L*C] -- the tensor HashEncoder.forward hands to the MLP -- so the
 * reference wrapper's [L,B,C] -> [B,L*C] permute copies (hashgrid.py:41,56; 134 MB read+write per 524 288-point
 * call) disappear.  Same arithmetic as the reference-layout ops. */
int ac_hash_encode_forward_pm(const float *inputs, const float *embeddings, const int32_t *offsets,
                              float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                              uint32_t H, int calc_grad_inputs, float *dy_dx, void *stream);
int ac_hash_encode_backward_pm(const float *grad, const float *inputs, const int32_t *offsets,
                               float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, int calc_grad_inputs, const float *dy_dx, float *grad_inputs,
                               void *stream);
/* Per-level scale exp2f(l*S)*H-1 as evaluated ON THE DEVICE (hashencoder.cu:121); scales [L]. */
int ac_hash_level_scales(float *scales, uint32_t L, float S, uint32_t H, void *stream);

/* --------------------------------------------------------------------------------------
 * Spherical-harmonics direction encoder.  Replaces
 *   sh_encode_forward  (encoder/shencoder/src/shencoder.h:11, shencoder.cu:387-393)
 *   sh_encode_backward (encoder/shencoder/src/shencoder.h:14, shencoder.cu:395-400)
 * inputs [B,3]; outputs [B,degree^2]; dy_dx [B,3,degree^2]; degree in 1..8.
 * ------------------------------------------------------------------------------------ */
int ac_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t degree,
                         int calc_grad_inputs, float *dy_dx, void *stream);
int ac_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t degree,
                          const float *dy_dx, float *grad_inputs, void *stream);

/* --------------------------------------------------------------------------------------
 * Instant-NSR model (models/instant_nsr.py:478-726): hash grid (L=16, C=2, D=3) +
 * SDF MLP 35->64->16 (softplus beta=100) + colour MLP 21->64->64->3 (relu, sigmoid)
 * + the single NeuS variance.
 * ------------------------------------------------------------------------------------ */
#define AC_NSR_MLP_BLOB_FLOATS 10848

/* Folds weight-norm (w = g*v/||v||_row, models/instant_nsr.py:555-556,585-586) and packs
 * the five layers into the blob the render kernels stage in shared memory.  Pointers are
 * the state-dict tensors: sdf_net.{0,1}.{weight_g,weight_v,bias},
 * color_net.{0,1,2}.{weight_g,weight_v}.  blob [AC_NSR_MLP_BLOB_FLOATS]. */
int ac_nsr_pack_mlp(const float *sdf0_g, const float *sdf0_v, const float *sdf0_b,
                    const float *sdf1_g, const float *sdf1_v, const float *sdf1_b,
                    const float *col0_g, const float *col0_v, const float *col1_g, const float *col1_v,
                    const float *col2_g, const float *col2_v, float *blob, void *stream);

typedef struct ac_nsr_model {
    const float *embeddings;   /* [n_entries,2] encoder.embeddings                              */
    const int32_t *offsets;    /* [17] encoder.offsets (device)                                 */
    const float *mlp_blob;     /* [AC_NSR_MLP_BLOB_FLOATS] from ac_nsr_pack_mlp                  */
    const float *variance;     /* [1] deviation_net.variance (device scalar)                    */
    float log2_per_level_scale;/* S (hashgrid.py:28)                                            */
    uint32_t base_resolution;  /* H = 16                                                        */
} ac_nsr_model;

/* NeRFNetwork.forward_sdf (models/instant_nsr.py:627-642) on a flat list of points.
 * x [B,3] in [-bound,bound] -> out [B,16] (col 0 sdf, cols 1..15 geometry features). */
int ac_nsr_forward_sdf(const ac_nsr_model *model, const float *x, float *out, uint32_t B, float bound,
                       void *stream);
/* Training: backward of ac_nsr_forward_sdf.  grad_out [B,16] = dL/d(out).  Accumulates the
 * hash-table gradient into grad_table [n_entries,2] (caller zeroes; fp32 reductions as in
 * kernel_grid_backward, hashencoder.cu:223-308) and writes the per-point layer terms from which
 * the host forms the weight gradients with plain GEMMs:
 *   delta_a [64,B] = dL/d(hidden pre-activation), hidden [64,B] = softplus output, feats [36,B] = the layer's
 *   input rows (x, y, z, 32 hash features, 1), fp32, unit-major (so the kernel's stores are coalesced and the host
 *   GEMMs read K-contiguous operands)
 *   => [dW0 | db0] = delta_a feats^T  ([64,36]: the ones row yields the bias gradient), dW1 = (hidden grad_out)^T,
 *   db1 = column sums of grad_out. */
int ac_nsr_sdf_backward(const ac_nsr_model *model, const float *x, const float *grad_out, uint32_t B,
                        float bound, float *grad_table, float *delta_a, float *hidden, float *feats,
                        void *stream);
/* Same backward with the weight gradients reduced in the kernel (tcgen05, accumulators in TMEM across the persistent
 * loop) instead of written out as per-point terms: grad_w0b [64,36] += s_d * [dW0 | db0], grad_w1 [16,64] += s_g * dW1
 * (both accumulated into; caller zeroes), grad_table as above.  scales: DEVICE pointer to (s_d, s_g), powers of two that
 * bring delta and grad_out into the fp16 range of the tensor-core operands: any s_d <= 30000 / (max|grad_out| *
 * max_j sum_o |W1[o][j]|) and s_g <= 30000 / max|grad_out|; the caller divides the results by them. */
int ac_nsr_sdf_backward_fused(const ac_nsr_model *model, const float *x, const float *grad_out, uint32_t B, float bound,
                              const float *scales, float *grad_table, float *grad_w0b, float *grad_w1, void *stream);
/* The training path's finite-difference stencil in one call each way (models/instant_nsr.py:210-215 + :683-704): P [M,3]
 * section points (already clamped to the bound).  Forward: out_centre [M,16] = forward_sdf(P), out_fd [6,M] = the signed
 * distance at clamp(P +- eps e_k) in the order (+x, -x, +y, -y, +z, -z); the 7 M points are generated inside the kernel
 * (no [7M,3] point list, no 15 unused outputs per neighbour).  Backward: grad_centre [M,16], grad_fd [6,M] -> the same
 * outputs as ac_nsr_sdf_backward_fused. */
int ac_nsr_forward_sdf_stencil(const ac_nsr_model *model, const float *P, uint32_t M, float bound, float eps,
                               float *out_centre, float *out_fd, void *stream);
int ac_nsr_sdf_backward_stencil(const ac_nsr_model *model, const float *P, uint32_t M, float bound, float eps,
                                const float *grad_centre, const float *grad_fd, const float *scales, float *grad_table,
                                float *grad_w0b, float *grad_w1, void *stream);
/* The same two backward entry points with a caller-provided workspace of ac_nsr_sdf_backward_workspace_bytes(points) bytes
 * (points = B, or 7 M for the stencil): the work is then split into a gather + tensor-core kernel that leaves d loss /
 * d(features) of every point in the workspace and a full-occupancy kernel that scatters it into grad_table -- faster than the
 * fused kernel, whose reductions and gathers wait on each other.  workspace = NULL (or too small) runs the fused kernel. */
uint64_t ac_nsr_sdf_backward_workspace_bytes(uint32_t n_points);
int ac_nsr_sdf_backward_fused_ws(const ac_nsr_model *model, const float *x, const float *grad_out, uint32_t B, float bound,
                                 const float *scales, float *grad_table, float *grad_w0b, float *grad_w1, void *workspace,
                                 uint64_t workspace_bytes, void *stream);
int ac_nsr_sdf_backward_stencil_ws(const ac_nsr_model *model, const float *P, uint32_t M, float bound, float eps,
                                   const float *grad_centre, const float *grad_fd, const float *scales, float *grad_table,
                                   float *grad_w0b, float *grad_w1, void *workspace, uint64_t workspace_bytes,
                                   const void *feature_cache, void *stream);
/* Feature cache between the stencil forward and its backward: ac_nsr_forward_sdf_stencil_cache also writes every 128-point
 * tile's encoded operand (32 hash features as fp16 hi | lo, 16 KB per tile; ac_nsr_sdf_feature_cache_bytes(7 M) bytes, 16-byte
 * aligned) and ac_nsr_sdf_backward_stencil_ws(feature_cache = that buffer, same P / M / model) loads it back instead of
 * repeating the 128 table gathers per point.  feature_cache = NULL recomputes the encoding. */
uint64_t ac_nsr_sdf_feature_cache_bytes(uint32_t n_points);
int ac_nsr_forward_sdf_stencil_cache(const ac_nsr_model *model, const float *P, uint32_t M, float bound, float eps,
                                     float *out_centre, float *out_fd, void *feature_cache, uint64_t cache_bytes, void *stream);
/* --------------------------------------------------------------------------------------
 * The differentiable half of NeRFRenderer.run on the training path, after the SDF stencil
 * (models/instant_nsr.py:210-299; driven by stylize.py:153-193): finite-difference normal, colour MLP, NeuS alpha,
 * transmittance scan, compositing, eikonal -- forward and backward as one launch each.
 *   n_rays rays x n_samples sorted depths (z_vals [n, T], T <= 128); M = n * T samples, m = ray * T + k.
 *   points [M,3]: section points clamp(o + d * z_mid) (ac_nsr_section_points); centre [M,16], fd [6,M]: outputs of
 *   ac_nsr_forward_sdf_stencil on `points`; num_steps: the coarse count (last interval = (far - near) / num_steps, :160,191).
 * Forward writes rgb [n,3] (over bg_color [n,3], NULL = white), depth [n], weight_sum [n], normal [n,3],
 *   eik_partial [n,2] (workspace), per-sample weights [M], pts_color [M,3], pts_alpha [M] (all required: the backward
 *   reads them) and eik_out[0] = sum(m (|g| - 1)^2) / (sum(m) + 1e-5) over the launch (:266-272), eik_out[1] = sum(m).
 * Backward: upstream gradients g_rgb [n,3] (required), g_weight_sum [n], g_normal [n,3], g_depth [n] (optional), g_eikonal
 *   (optional DEVICE scalar); or the trainer's opacity term fused: wsum_gt [n] (optional) adds opacity_weight *
 *   d/d ws mean(smooth_l1(clamp(ws,0,1), clamp(wsum_gt,0,1))) (stylize.py:187-193).
 *   Writes g_centre [M,16] and g_fd [6,M] (the inputs of ac_nsr_sdf_backward_stencil), adds d loss / d variance to
 *   *g_variance (optional), and writes the fp16 per-sample terms of the colour-MLP weight gradients, sample-contiguous:
 *     terms_a [136, ld]: rows 0..63 d/d a2 (layer-1 pre-activation), 64..127 d/d a1, 128..130 d/d z2, all times `scale`
 *     terms_b [160, ld]: rows 0..63 h1, 64..84 the layer-0 input (x, n, 15 features), 96..159 h2
 *   (rows 131..135 / 85..95 are never written: allocate zeroed), so that ONE ac_sd_gemm_f16(A = terms_a, W = terms_b,
 *   M = 136, N = 160, K = M samples) yields scale * [dC1 at [0:64, 0:64], dC0 at [64:128, 64:85], dC2 at [128:131, 96:160]].
 *   scale: DEVICE scalar, a power of two with scale * max|g_rgb| <= ~256 (ac_absmax_scale(g_rgb, 3 n, 256, scale)).
 * ------------------------------------------------------------------------------------ */
typedef struct {
  const float *rays_o, *rays_d, *z_vals, *points, *centre, *fd, *bg_color;
  uint32_t n_rays, n_samples, num_steps;
  float bound, eps, cos_anneal_ratio;
  float *rgb, *depth, *weight_sum, *normal, *eik_partial, *weights, *pts_color, *pts_alpha;
} ac_nsr_shade_args;
typedef struct {
  const float *g_rgb, *g_weight_sum, *g_normal, *g_depth, *g_eikonal, *wsum_gt;
  float opacity_weight;
  const float *eik_out, *scale;
  float *g_centre, *g_fd, *g_variance;
  void *terms_a, *terms_b;
  float *g_b1;           /* optional [16]: += column sums of g_centre (the six g_fd rows cancel) = d loss / d (sdf layer-1 bias) */
  float *opacity_loss;   /* optional DEVICE scalar: += the value of the fused opacity term (needs wsum_gt) */
  uint64_t terms_ld;     /* row stride of terms_a / terms_b in halves: >= M, a multiple of 8 (the GEMM's lda / ldw) */
} ac_nsr_shade_grads;
int ac_nsr_shade_forward(const ac_nsr_model *model, const ac_nsr_shade_args *args, float *eik_out, void *stream);
int ac_nsr_shade_backward(const ac_nsr_model *model, const ac_nsr_shade_args *args, const ac_nsr_shade_grads *grads,
                          void *stream);
/* *scale_out = 2^floor(log2(target / max|x|)) over n floats (1 when x == 0): operand scale of the fp16 backward paths. */
int ac_absmax_scale(const float *x, uint32_t n, float target, float *scale_out, void *stream);
/* points [n*T,3] = clamp(o + d * z_mid, +-bound), z_mid = z + (z_next - z) / 2, last sample at its own depth (:186-206). */
int ac_nsr_section_points(const float *rays_o, const float *rays_d, const float *z_vals, uint32_t n_rays,
                          uint32_t n_samples, float bound, float *points, void *stream);
/* Stage operators of the warped path (points travel through the SMPL warp between them, models/instant_nsr.py:155-172,461-475):
 * ac_nsr_ray_points: points [n*T,3] = o + d z (clamped to +-bound when bound > 0); z == NULL generates the coarse depths
 *   near + (far - near) * linspace(0, 1, T) from near_far [n,2] and also writes them to z_out [n,T].
 * ac_nsr_merge_gather: out [n,T+16] = cat(sdf [n,T], s_new [n,16]) gathered by the merge permutation `order` (:466-470). */
int ac_nsr_ray_points(const float *rays_o, const float *rays_d, const float *z, const float *near_far, uint32_t n_rays,
                      uint32_t n_samples, float bound, float *z_out, float *points, void *stream);
int ac_nsr_merge_gather(const float *sdf, const float *s_new, const int32_t *order, uint32_t n_rays, uint32_t T, float *out,
                        void *stream);
/* x = clamp(x, -bound, bound) in place over n floats (warped points, :172); sdf [B] = column 0 of forward_sdf's [B,16]. */
int ac_clamp_inplace(float *x, uint64_t n, float bound, void *stream);
int ac_nsr_take_sdf(const float *out16, uint64_t B, float *sdf, void *stream);
/* Backward of the weight-norm fold W = g * v / |v|_row (models/instant_nsr.py:555-556,585-586) for up to 5 layers in one
 * launch: dv += ..., dg += ... from dW [rows, ldw] * dW_scale (gradients of the folded weights). */
typedef struct {
  const float *dW, *v, *g;
  float *dv, *dg;
  int rows, cols, ldw;
  const float *scale;    /* optional DEVICE scalar: dW holds scale * gradient (the fp16 operand scales of the backward kernels) */
  float *db;             /* optional [rows]: += dW[:, db_col] / scale (a bias gradient kept in a column of the same accumulator) */
  int db_col;
} ac_weight_norm_layer;
int ac_nsr_weight_norm_backward(const ac_weight_norm_layer *layers, uint32_t n_layers, void *stream);
/* scales[0..1] = (s_d, s_g) of ac_nsr_sdf_backward_stencil / _fused from the gradients themselves (scales: 3 floats, the
 * third is scratch = max|g|); no host synchronisation. */
int ac_nsr_sdf_backward_scales(const ac_nsr_model *model, const float *g_centre, const float *g_fd, uint32_t M,
                               float *scales, void *stream);
/* out[i] = uniform [0,1) from a counter-based generator keyed by (seed, i): the per-sample jitter of the training path
 * (models/instant_nsr.py:161-162 draws torch.rand).  ac_zero: cudaMemsetAsync on the stream. */
int ac_fill_uniform(float *out, uint64_t n, uint64_t seed, void *stream);
int ac_zero(void *p, uint64_t bytes, void *stream);

/* NeRFNetwork.forward_color (models/instant_nsr.py:644-663, use_viewdirs=False):
 * x [B,3], normal [B,3], geo_feat [B,15] -> rgb [B,3]. */
/* ac_nsr_forward_color_bias: the same with a per-point layer-0 bias [B,64] (use_viewdirs, see ac_nsr_viewdir_bias). */
int ac_nsr_forward_color(const ac_nsr_model *model, const float *x, const float *normal,
                         const float *geo_feat, float *rgb, uint32_t B, void *stream);
int ac_nsr_forward_color_bias(const ac_nsr_model *model, const float *x, const float *normal, const float *geo_feat,
                              const float *c0_bias, float *rgb, uint32_t B, void *stream);
/* NeRFNetwork.gradient / finite_difference_normals_approximator (:683-704): [B,3] -> [B,3]. */
int ac_nsr_fd_gradient(const ac_nsr_model *model, const float *x, float *grad, uint32_t B, float bound,
                       float epsilon, void *stream);

/* The fused render core: NeRFRenderer.run (models/instant_nsr.py:133-299, render_can=True)
 * for n_rays rays in one launch -- near/far, coarse samples, SDF queries, upsample_steps/16
 * importance rounds (up_sample :410-459, sample_pdf :21-55, cat_z_vals :461-475), section
 * mid-points, finite-difference normals, colour, NeuS alpha and front-to-back compositing.
 *   rays_o, rays_d [n_rays,3].
 *   bg_color: NULL (white, bg=1) or [n_rays,3].
 *   jitter:   NULL (eval) or [n_rays,num_steps] uniform [0,1) replacing torch.rand (:162).
 *   alpha_mask: NULL or [n_rays, num_steps+upsample_steps] (the warp path's mask, :245-248).
 * Outputs (any of the per-sample pointers may be NULL to skip the store):
 *   rgb [n,3], depth [n], weight_sum [n], normal [n,3]                (32 B per ray)
 *   weights [n,T], pts_color [n,T,3], pts_alpha [n,T], z_vals [n,T]   (T = num_steps+upsample_steps)
 *   eikonal [ceil(n_rays/eikonal_segment)]: sum(m*(|g|-1)^2)/(sum(m)+1e-5) (:265-272) over each
 *     consecutive segment of `eikonal_segment` rays (0 = one segment = the whole launch).  The
 *     reference's driver renders 4096-ray batches and adds their means (render_utils.py:556-575);
 *     one launch with eikonal_segment = rays_per_batch reproduces that without 16 launches.
 * Sampling only: with rgb == NULL the launch stops after the importance rounds and writes just z_vals [n,T]
 *   (the training path needs the sample depths, not the image, models/instant_nsr.py:175-185) -- 112 of the 1008
 *   SDF evaluations per ray.  depth / weight_sum / normal / eikonal are then not written.
 * workspace: >= ac_nsr_render_workspace_bytes(n_rays) bytes, 16 B aligned.
 * Constraints: num_steps in [2,128], upsample_steps a multiple of 16, T <= 128. */
typedef struct ac_nsr_render_args {
    const float *rays_o, *rays_d;
    const float *bg_color, *jitter, *alpha_mask;
    /* Optional stage inputs (all NULL for the plain canonical render).  When the SMPL warp is active
     * (render_can=False) the host runs sampling and warping as separate launches and hands the fused
     * kernel its results: z_in [n,T] sorted sample depths (skips sampling), pts_in [n,T,3] the
     * (warped) section points at which the network is evaluated, near_far_in [n,2]. */
    const float *z_in, *pts_in, *near_far_in;
    uint32_t n_rays, num_steps, upsample_steps, eikonal_segment;
    float bound, cos_anneal_ratio, normal_epsilon_ratio;
    float *rgb, *depth, *weight_sum, *normal;
    float *weights, *pts_color, *pts_alpha, *z_vals;
    float *eikonal;
    void *workspace;
    uint64_t workspace_bytes;
    /* use_viewdirs=True (models/instant_nsr.py:564-569,646-650): NULL, or [n_rays,64] = the contribution of the ray direction's
     * 16 SH coefficients to colour layer 0 (ac_nsr_viewdir_bias); the packed model then holds the other 21 columns. */
    const float *c0_ray_bias;
    /* non-zero: skip the colour network (rgb / pts_color are then meaningless; depth, weight_sum, normal, weights, alpha,
     * eikonal are unchanged) -- the opacity target of the trainer's frozen net only needs weight_sum (stylize.py:177-181). */
    uint32_t opacity_only;
    /* non-zero (with alpha_mask): a block of 32 consecutive samples whose mask is all zero on the four rays evaluated
     * together is not evaluated at all -- its alpha is zero whatever the networks say (models/instant_nsr.py:245-248), so rgb,
     * depth, weight_sum, normal, weights and alpha are bit-identical; pts_color of such samples is 0 and they drop out of the
     * eikonal statistic (which the inference entry point, render_warp.py:88, discards). */
    uint32_t skip_masked;
} ac_nsr_render_args;

uint64_t ac_nsr_render_workspace_bytes(uint32_t n_rays);
/* out [n,64] = sh [n,16] w_sh^T, w_sh [64,16] = the SH columns of the (weight-norm folded) colour layer 0; sh from
 * ac_sh_encode_forward(dirs, degree 4). */
int ac_nsr_viewdir_bias(const float *sh, const float *w_sh, uint32_t n, float *out, void *stream);
int ac_nsr_render(const ac_nsr_model *model, const ac_nsr_render_args *args, void *stream);

/* Stage-level entry points used by the parity tests to pin the integer paths with
 * identical inputs (same device functions the fused kernel runs):
 * one importance round on given (z, sdf): z [n,T], sdf [n,T] -> z_new [n,16],
 * bins [n,16,2] (below, above) int32, merged z_out [n,T+16], order [n,T+16] int32
 * (source index into cat([z, z_new])).  alpha_out (optional) [n,T-1] receives the section
 * alphas; alpha_in (optional) [n,T-1] REPLACES them before the pdf/cdf/search, so the
 * integer path can be compared on bit-identical weights. */
int ac_nsr_debug_upsample(const float *rays_o, const float *rays_d, const float *z, const float *sdf,
                          uint32_t n_rays, uint32_t T, float inv_s, const float *alpha_in, float *alpha_out,
                          float *z_new, int32_t *bins, float *z_out, int32_t *order, void *stream);
/* The same round as a product operator (up_sample + cat_z_vals, models/instant_nsr.py:410-475) for the warped path, whose
 * signed distances between rounds come from warped / un-warped points evaluated by the caller (:166-172, :464-469). */
int ac_nsr_upsample_round(const float *rays_o, const float *rays_d, const float *z, const float *sdf, uint32_t n_rays,
                          uint32_t T, float inv_s, float *z_new, int32_t *bins, float *z_out, int32_t *order, void *stream);

/* --------------------------------------------------------------------------------------
 * SMPL-guided inverse warp (utils/ray_utils.py).
 * ac_warp_prepare_mesh: verts [n_verts,3], faces [n_faces,face_stride] int32 (first 3 columns are
 *   vertex ids, as read_obj returns 6 columns) -> opaque mesh records (ac_warp_mesh_bytes bytes).
 * ac_warp_samples_to_canonical = warp_samples_to_canonical (utils/ray_utils.py:62-90):
 *   pts [n,3] -> can_pts [n,3] = (sum_k b_k T[v_k])^-1 (p,1) (xyz, un-normalised), mask [n] =
 *   (closest squared distance < threshold) as 0/1 floats; optional closest [n,3], face_id [n],
 *   dist2 [n].  T [n_T,4,4] row-major fp32, last rows (0,0,0,c).
 * ac_mesh_guided_near_far = geometry_guided_near_far_torch (utils/ray_utils.py:277-294) with the
 *   cube fallback of models/instant_nsr.py:147-153: near_far [n_rays,2].
 * ------------------------------------------------------------------------------------ */
uint64_t ac_warp_mesh_bytes(uint32_t n_faces);
int ac_warp_prepare_mesh(const float *verts, const int32_t *faces, uint32_t face_stride, uint32_t n_faces,
                         void *mesh, void *stream);
int ac_warp_samples_to_canonical(const float *pts, uint32_t n_pts, const void *mesh, uint32_t n_faces,
                                 const float *T, float threshold, float *can_pts, float *mask,
                                 float *closest, int32_t *face_id, float *dist2, void *stream);
/* The same query for pts [n_rays, n_samples, 3] (consecutive samples of a ray are neighbours): a thread walks 8 samples and
 * seeds each search with the triangle-inequality bound from the previous one.  Bit-identical results, fewer boxes opened. */
int ac_warp_samples_to_canonical_rays(const float *pts, uint32_t n_rays, uint32_t n_samples, const void *mesh,
                                      uint32_t n_faces, const float *T, float threshold, float *can_pts, float *mask,
                                      float *closest, int32_t *face_id, float *dist2, void *stream);
/* Same query with the points visited in the caller's `order` (a permutation of 0..n-1, or NULL): results are
 * written at the original indices.  ac_warp_query_keys gives 30-bit Morton keys (inside the mesh's bounding box
 * grown by `margin`) whose argsort is the order that keeps the 32 queries of a warp spatially adjacent. */
int ac_warp_samples_to_canonical_ordered(const float *pts, const int32_t *order, uint32_t n_pts, const void *mesh,
                                         uint32_t n_faces, const float *T, float threshold, float *can_pts,
                                         float *mask, float *closest, int32_t *face_id, float *dist2, void *stream);
/* The section-point warp of the render path (models/instant_nsr.py:198-203,245-248): the caller multiplies alpha by `mask`,
 * so a sample with no triangle within sqrt(threshold) needs no closest point.  The search is bounded by the threshold;
 * such samples come back with mask 0 and can_pts = pts (a finite stand-in that never reaches the image); samples inside
 * the threshold get exactly what ac_warp_samples_to_canonical_ordered returns. */
int ac_warp_samples_to_canonical_masked(const float *pts, const int32_t *order, uint32_t n_pts, const void *mesh,
                                        uint32_t n_faces, const float *T, float threshold, float *can_pts,
                                        float *mask, void *stream);
int ac_warp_query_keys(const float *pts, uint32_t n_pts, const void *mesh, uint32_t n_faces, float margin,
                       int32_t *keys, void *stream);
int ac_mesh_guided_near_far(const float *rays_o, const float *rays_d, uint32_t n_rays, const float *verts,
                            uint32_t n_verts, float radius, float bound, float *near_far, void *stream);

/* --------------------------------------------------------------------------------------
 * Occupancy-grid ray marching + compositing: the six ops of the reference's `_raymarching` extension
 * (raymarching/src/raymarching.h:8-16, bindings.cpp:5-11), same argument order and layouts, fp32:
 *   march_rays_train              raymarching.cu:56-222   rays [N,3] = (ray id, first sample, #samples), counter [2]
 *   composite_rays_train_forward  raymarching.cu:232-301  "sigmas" are alphas (NeuS variant)
 *   composite_rays_train_backward raymarching.cu:315-391
 *   march_rays / composite_rays / compact_rays            raymarching.cu:497-747 (inference, alive-ray compaction)
 * Dead code in the reference (nothing imports the module); provided for API completeness.
 * ------------------------------------------------------------------------------------ */
int ac_march_rays_train(const float *rays_o, const float *rays_d, const float *grid, float mean_density,
                        int iter_density, float bound, uint32_t N, uint32_t H, uint32_t M, float *xyzs, float *dirs,
                        float *deltas, int32_t *rays, int32_t *counter, uint32_t perturb, void *stream);
int ac_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                    float bound, uint32_t M, uint32_t N, float *weights_sum, float *image, void *stream);
int ac_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                     const float *rgbs, const float *deltas, const int32_t *rays,
                                     const float *weights_sum, const float *image, float bound, uint32_t M, uint32_t N,
                                     float *grad_sigmas, float *grad_rgbs, void *stream);
int ac_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t, const float *rays_o,
                  const float *rays_d, float bound, uint32_t H, const float *grid, float mean_density,
                  const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas, uint32_t perturb,
                  void *stream);
int ac_composite_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, float *rays_t, const float *sigmas,
                      const float *rgbs, const float *normals, const float *deltas, float *weights_sum, float *depth,
                      float *image, float *normal_map, void *stream);
int ac_compact_rays(uint32_t n_alive, int32_t *rays_alive, const int32_t *rays_alive_old, float *rays_t,
                    const float *rays_t_old, int32_t *alive_counter, void *stream);

/* --------------------------------------------------------------------------------------
 * The stages either side of the render core.
 * ac_gen_rays: per-pixel rays generated on the device, row-major [H*W,3] float32.
 *   c2w: HOST pointer to the 4x4 row-major float64 camera-to-world matrix (read at call time, passed by value).
 *   Pixel coordinate of column i is x0 + x_step*i, of row j is y0 + y_step*j.
 *   convention 0 = cap2rays / shot_rays (utils/render_utils.py:363-376, utils/ray_utils.py:25-37):
 *     float64 back-projection of (x, y, 1) through K^-1 and c2w, float32 world point minus the float32 camera
 *     centre, normalised;
 *   convention 1 = SMPLDataset.gen_rays_pose (utils/SMPLDataset.py:86-103): float32 p = ((x-cx)/fx, -(y-cy)/fy, -1)
 *     normalised and rotated by the pose.
 * ac_select_background = select_background (utils/render_utils.py:953-987), out [n_rays,3]:
 *   key%4: 0 white, 1 black, 2 gaussian grey N(0.5,0.1) clamped (counter-based generator keyed by `seed`; the
 *   reference draws from torch's CPU generator, so only the distribution is comparable), 3 chessboard (0.8/0.2,
 *   squares of side/10 pixels, n_rays must be a square) blurred with torchvision's GaussianBlur((5,9)) at `sigma`
 *   (the reference draws sigma ~ U(0.1, 2.0) per call; the caller draws it here).
 * ac_adam_step = torch.optim.Adam(lr, betas, eps; no weight decay, no amsgrad) (stylize.py:355-363) on one flat
 *   fp32 buffer of n elements (16 B aligned pointers), `step` >= 1 the 1-based step count, gradients multiplied by
 *   grad_scale first.  Entries whose gradient and both moments are zero are left untouched (their update is 0).
 * ac_sdf_grid_points: the lattice of extract_fields (models/instant_nsr.py:706-731) for slabs [i0, i0+ni) of the
 *   slowest axis: pts [(ni*res*res),3], index (i*res+j)*res+k <-> (X[i],Y[j],Z[k]), X = linspace(min, max, res).
 *   bound_min / bound_max are HOST pointers to 3 floats.  Feed pts to ac_nsr_forward_sdf.
 * ------------------------------------------------------------------------------------ */
int ac_gen_rays(const double *c2w, double fx, double fy, double cx, double cy, uint32_t W, uint32_t H, double x0,
                double x_step, double y0, double y_step, int convention, float *rays_o, float *rays_d, void *stream);
int ac_select_background(int key, uint32_t n_rays, uint64_t seed, float sigma, float *out, void *stream);
int ac_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, uint64_t n, float lr, float beta1,
                 float beta2, float eps, uint32_t step, float grad_scale, void *stream);
int ac_sdf_grid_points(const float *bound_min, const float *bound_max, uint32_t resolution, uint32_t i0, uint32_t ni,
                       float *pts, void *stream);
/* Iso-surface of a [res,res,res] lattice volume (index (i*res+j)*res+k) at `threshold` by marching tetrahedra --
 * stands in for mcubes.marching_cubes in extract_geometry (models/instant_nsr.py:733-752; PyMCubes is not in the
 * image).  Appends triangles through *counter (device uint64, caller zeroes): tri_pos [capacity,3,3] world positions,
 * tri_key [capacity,3] = (lower lattice id << 32 | higher lattice id) of the edge each vertex lies on (the caller
 * welds vertices by key).  *counter ends at the number of triangles FOUND; when it exceeds `capacity` only the first
 * `capacity` were stored -- re-run with room (capacity 0 = count only).  Triangle normals point towards increasing
 * field value.  bound_min / bound_max: HOST pointers to 3 floats. */
int ac_iso_surface(const float *volume, const float *bound_min, const float *bound_max, uint32_t resolution,
                   float threshold, float *tri_pos, int64_t *tri_key, uint64_t capacity, uint64_t *counter,
                   void *stream);

/* --------------------------------------------------------------------------------------
 * Stable-Diffusion UNet forward for the SDS step (reference: models/diffusion.py:121-132 calls diffusers'
 * UNet2DConditionModel under torch.no_grad(); third-party, parity unpinned).  Activations fp32 NHWC; GEMM operands
 * fp16 (`void *` = __half), fp32 accumulation in TMEM.
 * ac_sd_gemm_f16: C[M,N] = A[M,K] W[N,K]^T (+ bias[N]) (+ group_bias[m / rows_per_group][N]) (+ residual[M,N], row
 *   stride ldr); C fp32 or fp16 (out_f16) with row stride ldc.  lda/ldw and the batch strides of A/W are multiples of
 *   8 halves and the pointers 16 B aligned; M, N, K are arbitrary (tiles are zero-filled).  Batch index
 *   z in [0, batch_outer*batch_inner): operand offsets (z / batch_inner) * s?o + (z % batch_inner) * s?i (elements).
 * Producers of fp16 operands (all take fp32 inputs):
 *   ac_sd_group_norm_stats  NHWC [B,HW,C] -> stats [B,G,2] = (mean, rstd); sums_workspace: 2*B*G doubles.
 *   ac_sd_im2col_f16        NHWC [B,Hs,Ws,C] -> [B*Ho*Wo, k*k*Cp] (K index (ky*k+kx)*Cp + c, Cp = C rounded up to 8),
 *                           k in {1,3}, stride in {1,2}, zero pad `pad` top/left, optional nearest x2 up-sampling of the
 *                           source, optional GroupNorm(+SiLU) from `gn_stats` applied on the fly.
 *   ac_sd_layer_norm_f16, ac_sd_geglu_f16 (value * gelu(gate), x [M, 2*inner]), ac_sd_softmax_f16 (softmax(scale * s)
 *   over L columns, output row stride ld_out with zeroed padding), ac_sd_cast_f16.
 * ------------------------------------------------------------------------------------ */
int ac_sd_gemm_f16(const void *A, const void *W, const float *bias, const float *group_bias, int rows_per_group,
                   const float *residual, void *C, int out_f16, int M, int N, int K, int64_t lda, int64_t ldw,
                   int64_t ldc, int64_t ldr, int batch_outer, int batch_inner, int64_t sAo, int64_t sAi, int64_t sWo,
                   int64_t sWi, int64_t sCo, int64_t sCi, void *stream);
/* Fused attention: out[b, :, h*d:(h+1)*d] = softmax(scale * Q_bh K_bh^T) V_bh for every (batch, head), the score matrix
 * never leaves the SM (two passes over the key tiles: statistics, then P V with the accumulator in TMEM).
 * q [B*Lq, ld_q], k [B*Lk, ld_k] fp16 with head h at columns h*d; vt [B, heads*d, ld_vt] fp16 = V transposed (keys
 * contiguous, ld_vt >= Lk); out [B*Lq, ld_out] fp16.  d a multiple of 8, d <= 128; every ld a multiple of 8. */
int ac_sd_flash_attention_f16(const void *q, const void *k, const void *vt, void *out, int B, int heads, int Lq, int Lk,
                              int d, int64_t ld_q, int64_t ld_k, int64_t ld_vt, int64_t ld_out, float scale, void *stream);
/* 3x3 / stride 1 / zero-pad 1 convolution as an implicit GEMM (no im2col buffer): every k-tile's A operand is the
 * activation window of the CTA's 128 output pixels shifted by one of the 9 taps, fetched as a 4-D TMA box over
 * (C, W, H, B) whose out-of-image part the TMA unit zero-fills.  act fp16 NHWC [B,H,W,C] (C a multiple of 64; W divides
 * 128 and 128/W divides H, or H*W divides 128), w fp16 [N, 9*C] tap-major, out fp32 NHWC [B,H,W,N]
 * (+ bias[N], + group_bias[B,N], + residual[B,H,W,N]). */
int ac_sd_conv3x3_f16(const void *act, const void *w, const float *bias, const float *group_bias, const float *residual,
                      float *out, int B, int H, int W, int C, int N, void *stream);
int ac_sd_group_norm_stats(const float *x, int B, int HW, int C, int G, float eps, double *sums_workspace, float *stats,
                           void *stream);
int ac_sd_im2col_f16(const float *x, int B, int Hs, int Ws, int C, int ksize, int stride, int pad, int upsample2x,
                     int Ho, int Wo, const float *gn_stats, const float *gn_gamma, const float *gn_beta, int gn_groups,
                     int gn_silu, void *out, void *stream);
int ac_sd_layer_norm_f16(const float *x, int M, int C, const float *gamma, const float *beta, float eps, void *out,
                         void *stream);
int ac_sd_geglu_f16(const float *x, int64_t M, int inner, void *out, void *stream);
int ac_sd_softmax_f16(const float *scores, int64_t rows, int L, int64_t ld_in, int64_t ld_out, float scale, void *out,
                      void *stream);
int ac_sd_cast_f16(const float *x, int64_t n, void *out, void *stream);
/* Backward pieces of the VAE encoder of the SDS step (models/diffusion.py:304-312 encodes WITH gradient, :148 back-propagates
 * the latent gradient to the image; the weights are frozen, so only data gradients exist):
 *   ac_sd_group_norm_backward   x, dy NHWC [B,HW,C], stats [B,G,2] of the forward -> dx (+ add) as fp32 and / or fp16;
 *                               y = silu_act ? silu(gn(x)) : gn(x); sums_workspace: 2*B*G doubles.
 *   ac_sd_softmax_backward_f16  dS = P (dP - rowsum(P dP)) * scale; P fp16, dP fp32, both [rows, ld] -> dS fp16 [rows, ld].
 *   ac_sd_transpose_f16         fp16 [rows, cols] -> [cols, rows].
 *   ac_sd_conv_s2_dgrad_operand_f16  GEMM operand [B*H*W, 9*Np] of the input gradient of the 3x3 stride-2 convolution with
 *                               bottom/right zero padding (Downsample2D of the VAE) from dy NHWC [B,Ho,Wo,N]; Np = N rounded up to 8.
 * Input gradients of the stride-1 convolutions are ac_sd_conv3x3_f16 / ac_sd_gemm_f16 with the flipped, transposed weights. */
int ac_sd_group_norm_backward(const float *x, const float *dy, int B, int HW, int C, int G, const float *stats, const float *gamma,
                              const float *beta, int silu_act, const float *add, float *dx32, void *dx16, double *sums_workspace,
                              void *stream);
int ac_sd_softmax_backward_f16(const void *probs, const float *dprobs, int64_t rows, int L, int64_t ld, float scale, void *dscores,
                               void *stream);
int ac_sd_transpose_f16(const void *in, int rows, int cols, int64_t ld_in, void *out, int64_t ld_out, void *stream);
int ac_sd_conv_s2_dgrad_operand_f16(const float *dy, int B, int Ho, int Wo, int N, int H, int W, void *out, void *stream);

/* Unit test of the tensor-core layer in isolation: feats [128,32] fp32 x (sdf layer 0 feature
 * columns)^T -> out [128,64] pre-activations WITHOUT bias / xyz terms (3xTF32 tcgen05.mma). */
int ac_nsr_debug_tc_layer(const float *feats, const float *mlp_blob, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AVATARCRAFT_B200_H */
