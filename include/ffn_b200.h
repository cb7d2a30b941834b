/*
 * ffn_b200.h -- C ABI of libffn_b200.so: the B200 (sm_100a) volume-rendering hot path
 * that replaces, for matajoh/fourier_feature_nets,
 *
 *     RaySampler.sample            fourier_feature_nets/ray_sampler.py:359-403
 *  -> NeRF.forward                 fourier_feature_nets/nerf_model.py:86-124
 *     FourierFeatureMLP.forward    fourier_feature_nets/fourier_feature_models.py:57-78
 *  -> Raycaster.render             fourier_feature_nets/ray_caster.py:48-93
 *     calculate_blend_weights      fourier_feature_nets/utils.py:72-97
 *
 * The reference has no FFI of its own (pure Python); these entry points sit one level
 * below its Python seams and are what a ctypes binding in the
 * reference would call (INTEGRATION.md shows that binding).
 *
 * Conventions
 *  - plain pointers and sizes only; every data pointer is a DEVICE pointer on the
 *    current CUDA device, contiguous row-major float32, 16-byte aligned;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *    synchronises;
 *  - return 0 on success, non-zero on error; ffn_last_error() gives the message
 *    (thread-local);
 *  - the library owns only what lives inside an ffn_net_t handle.
 */
#ifndef FFN_B200_H_
#define FFN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFN_B200_VERSION 100

/* operand type fed to the tensor cores (accumulation, bias, activations, encodings,
 * heads and compositing are always float32) */
#define FFN_OPERAND_FP16 0
#define FFN_OPERAND_BF16 1
/* NeRF handles, inference entry points: every operand split into fp16 high + residual parts, three UMMAs per product
 * (x_hi w_hi + x_lo w_hi + x_hi w_lo, fp32 accumulate): fp32-grade pixels at ~1/6 of the fp16 throughput.  Training entry
 * points of such a handle run with fp16 operands. */
#define FFN_OPERAND_FP16X3 2

#define FFN_MAX_LAYERS 16
#define FFN_MAX_FREQS 10

typedef struct ffn_net ffn_net_t;

/* NeRF(num_layers, num_channels, max_log_scale_pos, num_freq_pos, max_log_scale_view,
 *      num_freq_view, skips, include_inputs)   -- nerf_model.py:12-75.
 * freq_pos / freq_view are the diagonal of the module's frozen pos_encoding /
 * view_encoding buffers (2**linspace(0, max_log_scale, F), nerf_model.py:77-84). */
typedef struct {
  int32_t num_layers;
  int32_t num_channels;      /* must be 256 (hidden_view = 128) */
  int32_t num_freq_pos;      /* <= 10 */
  int32_t num_freq_view;     /* <= 10 */
  int32_t include_inputs;
  int32_t num_skips;
  int32_t skips[FFN_MAX_LAYERS];
  float freq_pos[FFN_MAX_FREQS];
  float freq_view[FFN_MAX_FREQS];
  int32_t operand_dtype;     /* FFN_OPERAND_* */
} ffn_nerf_desc_t;

int ffn_version(void);
const char* ffn_last_error(void);

/* Build a handle for a NeRF of the given shape (allocates the packed-weight arena). */
int ffn_nerf_create(const ffn_nerf_desc_t* desc, ffn_net_t** out);

/* FourierFeatureMLP(num_inputs=3, num_outputs=4, a_values, b_values, layer_channels=[256]*n)
 * -- fourier_feature_models.py:13-55.  a_values (E) / b_values (3,E) are HOST pointers
 * (frozen buffers); NULL = the un-encoded MLP preset.  E <= 256. */
int ffn_ffmlp_create(int32_t num_hidden_layers, int32_t num_channels, int32_t embedding_size,
                     const float* a_values_host, const float* b_values_host,
                     int32_t operand_dtype, ffn_net_t** out);

void ffn_net_destroy(ffn_net_t* net);

/* Number of (weight, bias) pairs ffn_net_pack expects and their order:
 * NeRF: layers.0 .. layers.{L-1}, opacity_out, bottleneck, hidden_view, color_out.
 * FourierFeatureMLP: layers.0 .. layers.{n}. */
int ffn_net_num_linear(const ffn_net_t* net);

/* Re-pack torch-layout (out,in) float32 weights + biases (device pointers, given as HOST
 * arrays of pointers) into the tensor-core layout.  Call after every optimiser step.
 * NeRF handles remember the pointers: the first INFERENCE launch after a pack reads bottleneck / hidden_view once
 * more (stream-ordered, one small launch) to build the folded layer its program runs (bottleneck has no activation,
 * nerf_model.py:119-122, so bottleneck . hidden_view is one 128-wide layer).  The parameters must therefore stay
 * allocated, and unchanged unless ffn_net_pack is called again, until that launch -- true for a model's own
 * nn.Parameter storage.  FFN_FOLD=0 in the environment keeps the reference's layer structure. */
int ffn_net_pack(ffn_net_t* net, const float* const* weights, const float* const* biases,
                 void* stream);

/* model(positions[, views]) -> (N,4) raw [rgb | sigma]   (nerf_model.py:86, ray_caster.py:61-66) */
int ffn_mlp_forward(ffn_net_t* net, const float* positions, const float* views, int64_t n,
                    float* out4, void* stream);

/* Raycaster.render on materialised RaySamples (ray_caster.py:48-93):
 * positions (R,S,3), view_directions (R,S,3) or NULL, t_values (R,S)
 * -> color (R,3), alpha (R), depth (R) or NULL.  nan_flag: device int, OR-ed with 1 when a
 * NaN colour/opacity is produced (the reference's asserts at ray_caster.py:73-74).
 * ONE kernel launch for any num_samples >= 1 (ffn_render_samples / _rays / _rays_t): when 128 % num_samples != 0 rays
 * straddle the kernel's 128-sample tiles and are finished through per-ray partials in a scratch buffer owned by the
 * handle (grown on demand: 32 * ((S + 126) / 128 + 1) + 4 bytes per ray; one stream per handle at a time). */
int ffn_render_samples(ffn_net_t* net, const float* positions, const float* view_directions,
                       const float* t_values, int64_t num_rays, int32_t num_samples,
                       float* color, float* alpha, float* depth, int32_t* nan_flag, void* stream);

/* RaySampler.sample (no focus sampling) fused into the render:
 * starts (R,3), directions (R,3), near (R), far (R) [already annealed, ray_sampler.py:373-378],
 * lin = torch.linspace(0,1,S) (S floats), jitter (R,S) uniform draws or NULL.
 * stratified != 0 with jitter == NULL draws Philox(seed, ray_offset + ray, sample) in-kernel. */
int ffn_render_rays(ffn_net_t* net, const float* starts, const float* directions,
                    const float* near, const float* far, const float* lin, const float* jitter,
                    int32_t stratified, uint64_t seed, int64_t ray_offset, int64_t num_rays,
                    int32_t num_samples, float* color, float* alpha, float* depth,
                    float* t_values_out /* (R,S) or NULL */, int32_t* nan_flag, void* stream);

/* Compositing alone (ray_caster.py:67-93 + utils.py:72-97) for any S:
 * raw (R,S,4), t_values (R,S) -> color, alpha, depth (or NULL), weights (R,S) (or NULL). */
int ffn_composite(const float* raw, const float* t_values, int64_t num_rays, int32_t num_samples,
                  float* color, float* alpha, float* depth, float* weights, int32_t* nan_flag,
                  void* stream);

/* calculate_blend_weights(t_values (R,S), opacity (R,S)) -> weights (R,S)   (utils.py:72-97) */
int ffn_blend_weights(const float* t_values, const float* opacity, int64_t num_rays,
                      int32_t num_samples, float* weights, void* stream);

/* ---- hierarchical ("focus") sampling, ray_sampler.py:59-67,234-269,301-357,388-392, done per batch on
 * the GPU instead of a constructor-time pass that stores a (num_rays, S_c-1) CDF table on the host ---- */

/* Sorted t values (R,S) from the coarse model's raw outputs `raw` ((R, S_c, raw_stride), opacity logit in the last
 * channel; S_c = S - S/2) at t_c = near + lin_c (far - near).  near_u/far_u: the (possibly annealed) segment of
 * the S/2 uniform samples; lin_u = linspace(0,1,S/2), lin_c = lin_f = linspace(0,1,S_c); jitter_u (R,S/2) and
 * u_focus (R,S_c) are the reference's torch.rand draws (NULL: Philox when stratified, else no jitter / lin_f). */
int ffn_focus_t(const float* raw, int32_t raw_stride, const float* near, const float* far, const float* near_u,
                const float* far_u, const float* lin_c, const float* lin_u, const float* lin_f,
                const float* jitter_u, const float* u_focus, int32_t stratified, uint64_t seed,
                int64_t ray_offset, int64_t num_rays, int32_t num_samples, float* t_out, void* stream);

/* coarse sigma pass of `coarse` (fused MLP kernel, scratch inside the handle) + ffn_focus_t */
int ffn_focus_sample(ffn_net_t* coarse, const float* starts, const float* directions, const float* near,
                     const float* far, const float* near_u, const float* far_u, const float* lin_c,
                     const float* lin_u, const float* lin_f, const float* jitter_u, const float* u_focus,
                     int32_t stratified, uint64_t seed, int64_t ray_offset, int64_t num_rays,
                     int32_t num_samples, float* t_out, void* stream);

/* Raycaster.render on rays with explicit per-sample t (R,S): positions = starts + t * directions in-kernel */
int ffn_render_rays_t(ffn_net_t* net, const float* starts, const float* directions, const float* t_values,
                      int64_t num_rays, int32_t num_samples, float* color, float* alpha, float* depth,
                      int32_t* nan_flag, void* stream);

/* ---- ray tables on the device: CameraInfo.unproject/raycast (camera_info.py:66-74,99-109) and
 * RaySampler._near_far (ray_sampler.py:202-232) for every pixel of every camera.  unproj (C,16): row-major
 * inv(K~ . inv(E)) per camera and cam_pos (C,3) are DEVICE pointers; bounds_min/bounds_max are HOST pointers to
 * the 3 floats of bounds @ (-+.5,-+.5,-+.5,1).  Ray c*H*W + y*W + x is pixel (x,y) of camera c.  Outputs:
 * starts, directions (C*H*W,3), near_far (2,C*H*W) [near clamped to >= 0.1 on hits], valid (C*H*W) bytes. ---- */
int ffn_generate_rays(const float* unproj, const float* cam_pos, const float* bounds_min, const float* bounds_max,
                      int32_t num_cameras, int32_t width, int32_t height, float* starts, float* directions,
                      float* near_far, uint8_t* valid, void* stream);

/* Voxels.forward (voxels_model.py:35-45): trilinear grid_sample (border padding, align_corners=False) of
 * positions / scale + bias.  grid_channels_last: DEVICE (side,side,side,4) = voxels[0].permute(1,2,3,0);
 * bias4: HOST pointer to 4 floats; positions (n,3) and out4 (n,4) DEVICE. */
int ffn_voxels_forward(const float* grid_channels_last, const float* bias4, int32_t side, float scale,
                       const float* positions, int64_t n, float* out4, void* stream);

/* ---- training step (ray_caster.py:95-101,319-329): forward with saves, compositing backward, dgrad chain.
 * Weight gradients dW = dz^T x are plain GEMMs over the saved tensors and are left to the caller. ---- */

/* slot counts of the workspaces: save_h [n_save][M][256] bf16, save_mask [n_mask][M][8] u32,
 * dz_out [n_dz][M][256] bf16 (slot = forward MMA layer index: trunk 0..L-1, bottleneck L, hidden_view L+1) */
int ffn_train_slots(const ffn_net_t* net, int32_t* n_save, int32_t* n_mask, int32_t* n_dz);

/* transposed bf16 weight images for the dgrad chain (call after the weights changed, before backward) */
int ffn_net_pack_backward(ffn_net_t* net, const float* const* weights, void* stream);

/* Raycaster.render in training mode: inputs either materialised samples (positions, view_directions,
 * t_values; the ray pointers NULL) or rays (starts, directions, near, far, lin[, jitter]; sample pointers
 * NULL).  Also writes raw (M,4), t_out (R,S) [rays mode], save_h, save_mask, save_enc ([2][M][64]) -- both 16-bit saves in the
 * operand dtype of the net (fp16 by default, bf16 with FFN_OPERAND_BF16). */
int ffn_train_forward(ffn_net_t* net, const float* positions, const float* view_directions,
                      const float* t_values, const float* starts, const float* directions,
                      const float* near, const float* far, const float* lin, const float* jitter,
                      int32_t stratified, uint64_t seed, int64_t ray_offset, int64_t num_rays,
                      int32_t num_samples, float* color, float* alpha, float* depth, float* raw,
                      float* t_out, void* save_h, void* save_mask, void* save_enc, int32_t* nan_flag,
                      void* stream);

/* d(loss)/d(color (R,3), alpha (R) or NULL) -> d(loss)/d(raw (R,S,4))   (num_samples <= 256) */
int ffn_composite_backward(const float* raw, const float* t_values, int64_t num_rays, int32_t num_samples,
                           const float* grad_color, const float* grad_alpha, float* d_raw, void* stream);

/* dgrad chain: d_raw (M,4) + sign words -> dz_out */
int ffn_train_backward(ffn_net_t* net, const float* d_raw, const void* save_mask, int64_t num_points,
                       void* dz_out, void* stream);

/* Bias gradients of all layers at once: x = dz (num_slots, M, 256) bf16 -> out (num_slots, 256) fp32 column sums
 * (autograd of nn.Linear's bias, nerf_model.py:111-122). */
int ffn_colsum_bf16(const void* x, int32_t num_slots, int64_t num_points, float* out, void* stream);

/* Gradients of a 1..4-row head evaluated on CUDA cores (opacity_out nerf_model.py:118, color_out :123, final Linear
 * fourier_feature_models.py:77): out_w[o][c] = sum_m d_raw[m][first_head+o] * h[m][c] (h (M,256) bf16, or fp16 when
 * h_fp16 is set -- the dtype ffn_train_forward saved it in --, fp32 accumulate),
 * out_b[o] = sum_m d_raw[m][first_head+o];  out_w (num_heads, num_cols) keeps the first num_cols <= 256 columns
 * (color_out reads the 128 hidden_view channels), out_b (num_heads). */
int ffn_head_wgrad(const float* d_raw, int32_t first_head, int32_t num_heads, const void* h, int64_t num_points,
                   float* out_w, float* out_b, int32_t num_cols, int32_t h_fp16, void* stream);

/* Weight (and bias) gradients of every MMA layer in ONE launch: dW[out][in] += sum_rows dz[row][out] * x[row][in]
 * (autograd of nn.Linear inside Raycaster.fit, ray_caster.py:319-326; layers nerf_model.py:111-123,
 * fourier_feature_models.py:70-77).  Split-K tcgen05 GEMM over the saved tensors, accumulating with red.global.add
 * into fp32 destinations the CALLER HAS ZEROED.
 *   tensors[i]: [slots][rows][cols] row-major bf16 -- or fp16 when .fp16 is set (B operands only: converted to bf16
 *               in shared memory; dz, the A operand, must be bf16) --, cols a multiple of 64, the same
 *               `rows` for all, 16-byte aligned.
 *   jobs[j]   : A = n_mtiles (1|2) tiles of 128 columns of slot a_slot of tensors[a_tensor] from a_col0,
 *               B = n_cols (64..256, multiple of 64) columns of slot b_slot of tensors[b_tensor] from b_col0;
 *               dst[(128*mt + r) * dst_stride + dst_col0 + c] += D[mt][r][c]  for c < dst_cols, or, with colmap,
 *               dst[... + colmap[c]] for colmap[c] >= 0 (colmap: device int32[n_cols]);
 *               bias_dst (optional, device float[128*n_mtiles]) += column sums of the A tile. */
typedef struct {
  const void* ptr;
  int64_t rows;
  int32_t cols, slots, fp16;
} ffn_wgrad_tensor_t;
typedef struct {
  int32_t a_tensor, a_slot, a_col0, n_mtiles;
  int32_t b_tensor, b_slot, b_col0, n_cols;
  float* dst;
  int32_t dst_stride, dst_col0, dst_cols;
  const int32_t* colmap;
  float* bias_dst;
} ffn_wgrad_job_t;
int ffn_wgrad(const ffn_wgrad_tensor_t* tensors, int32_t n_tensors, const ffn_wgrad_job_t* jobs, int32_t n_jobs,
              void* stream);

/* Optimiser step of Raycaster.fit (ray_caster.py:327-329) in two launches: clip_grad_value_(clip_value) ->
 * clip_grad_norm_(max_norm) (both written back into grad, <= 0 disables) -> torch.optim.Adam update with L2
 * weight decay.  bias_correction{1,2} = 1 - beta{1,2}^step.  norm_scratch (device float[norm_scratch_floats], at least
 * 1 + sum ceil(numel / 2048)): [0] receives the squared total norm of the value-clipped gradients, the rest holds
 * per-block partial sums that are added in a fixed order (deterministic: data-parallel replicas stay bit-identical).
 * All tensors fp32, contiguous, on the current device. */
typedef struct {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} ffn_adam_tensor_t;
int ffn_clip_adam(const ffn_adam_tensor_t* tensors, int32_t n, float clip_value, float max_norm, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                  float bias_correction2, float* norm_scratch, int32_t norm_scratch_floats, void* stream);

/* Loss of Raycaster.fit and its gradient in one launch (ImageDataset.render/.loss, image_dataset.py:224-262):
 * loss = mean((colors[rays] - color)^2) + alpha_weight * mean((alphas[rays] - alpha)^2), ground-truth colour zeroed
 * where the ground-truth alpha is 0; gt_alphas == NULL drops the alpha term (and the zeroing).  rays: int64 indices
 * into the ground-truth tables.  loss: device float[1]; grad_color (R,3), grad_alpha (R) = d loss / d prediction. */
int ffn_mse_loss(const float* color, const float* alpha, const float* gt_colors, const float* gt_alphas,
                 const int64_t* rays, int64_t num_rays, float alpha_weight, float* loss, float* grad_color,
                 float* grad_alpha, void* stream);

/* ---- One optimisation step of Raycaster.fit (ray_caster.py:319-329) driven from C: the ~12 launches of a step
 * issued back to back without interpreter time in between.  NeRF nets.  The caller owns every buffer:
 *   weights/biases   the fp32 parameters in ffn_net_pack order
 *   flat_grad        one fp32 buffer receiving every gradient at weight_grad_offset[i] / bias_grad_offset[i]
 *                    (floats; parameter layout (out,in) row-major) -- all-reduce it between the two calls for
 *                    data-parallel training
 *   exp_avg, exp_avg_sq   Adam state, same layout as flat_grad, zero-initialised
 *   workspace        >= ffn_trainer_workspace_bytes(net, num_rays, num_samples) bytes, 256-byte aligned */
typedef struct ffn_trainer ffn_trainer_t;
typedef struct {
  int32_t num_linear;
  float* const* weights;
  float* const* biases;
  const int64_t* weight_grad_offset;
  const int64_t* bias_grad_offset;
  float* flat_grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t flat_floats;
} ffn_trainer_desc_t;
int ffn_trainer_create(ffn_net_t* net, const ffn_trainer_desc_t* desc, ffn_trainer_t** out);
void ffn_trainer_destroy(ffn_trainer_t* trainer);
int64_t ffn_trainer_workspace_bytes(const ffn_net_t* net, int64_t num_rays, int32_t num_samples);
/* forward (inputs as ffn_train_forward: samples OR rays) + ffn_mse_loss against gt tables indexed by rays + every
 * gradient into flat_grad.  loss: device float[1]. */
int ffn_trainer_backward(ffn_trainer_t* trainer, const float* positions, const float* view_directions,
                         const float* t_values, const float* starts, const float* directions, const float* near,
                         const float* far, const float* lin, const float* jitter, int32_t stratified, uint64_t seed,
                         int64_t num_rays, int32_t num_samples, const float* gt_colors, const float* gt_alphas,
                         const int64_t* rays, float alpha_weight, void* workspace, int64_t workspace_bytes, float* loss,
                         int32_t* nan_flag, void* stream);
/* Data-parallel training: the flat gradient buffer holds the SUM over `1 / grad_scale` ranks after the all-reduce;
 * ffn_trainer_update multiplies every gradient by grad_scale (default 1) before clipping, i.e. the mean costs no
 * extra pass over the buffer.  The reference is single-device (ray_caster.py:319-329 has no counterpart). */
int ffn_trainer_set_grad_scale(ffn_trainer_t* trainer, float grad_scale);
/* ffn_clip_adam over all parameters + ffn_net_pack */
int ffn_trainer_update(ffn_trainer_t* trainer, float clip_value, float max_norm, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float bias_correction1, float bias_correction2,
                       float* norm_scratch, int32_t norm_scratch_floats, void* stream);

/* Debug: dump the float32 post-activation output of MMA layer `layer` (row-major (N,256))
 * for the first `n` points.  Used by the bring-up tests only. */
int ffn_debug_layer(ffn_net_t* net, const float* positions, const float* views, int64_t n,
                    int32_t layer, float* out256, void* stream);

/* Debug: with FFN_STATS=1 in the environment the render kernel accumulates 32 cycle counters: [0..3] UMMA-issuer
 * warps {total, waiting for the epilogue, waiting for weights, issuer count}, [4..7] epilogue warp 4 of every
 * CTA {waiting for the accumulator, converting, tile front (inputs + encoding), compositing}, [8..31] the
 * converting time split by layer.  This reads (after a device sync) and clears them. */
int ffn_debug_stats(ffn_net_t* net, uint64_t* out32);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
int64_t ffn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FFN_B200_H_ */
