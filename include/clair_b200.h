/*
 * clair_b200.h — C ABI of the B200-native radiometric hot path (libclair_b200.so).
 *
 * The reference (samivout/clair-torch) is pure Python/torch and has no FFI layer; the boundary it
 * exposes for this path is its Python API.  Each entry point below replaces the
 * per-pixel arithmetic of one reference function and is what a reference-side binding (ctypes, see
 * INTEGRATION.md) would call.  Citations are file:line in the reference repository.
 *
 * Conventions
 *   - every pointer named *_dev is DEVICE memory (fp32 unless stated), contiguous; image stacks are
 *     (N, C, H, W) with W fastest, exactly as `datasets/collate.py:8-43` hands them over;
 *   - every pointer named *_host is HOST memory read synchronously during the call;
 *   - `plane` = H*W of the tensor handed over; `stream` is a cudaStream_t (NULL = default stream);
 *   - all launches are asynchronous and stream-ordered; no call synchronises the device;
 *   - return value: 0 on success, a negative CLAIR_E_* code for a rejected argument, a positive
 *     cudaError_t if the CUDA runtime reported an error.  Nothing throws.  `clair_last_error()` returns
 *     a thread-local description of the last non-zero return.
 *   - `curve_row_base_host[c]`: the reference's LINEAR mode reads row `k mod C` of the (C, L) table for
 *     the element with flat NCHW index k (models/base.py:173-176).  For a whole image the
 *     row of element (c, p) is `(c*plane + p) mod C`; for a spatial shard the caller passes
 *     `curve_row_base_host[c] = (flat index of the shard's element (c, 0) in the full frame) mod C`.
 *     NULL means "whole image" (`(c*plane) mod C`).
 */
#ifndef CLAIR_B200_H
#define CLAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLAIR_API __attribute__((visibility("default")))
#else
#define CLAIR_API
#endif

#define CLAIR_ABI_VERSION 4
#define CLAIR_MAX_FRAMES 64     /* exposure frames per batch (exposure times travel as kernel arguments) */
#define CLAIR_MAX_CHANNELS 8
#define CLAIR_MAX_LUT 1024      /* ICRF samples per channel (reference default 256) */
#define CLAIR_MAX_PAIRS 2016    /* exposure pairs per batch = 64*63/2 */

#define CLAIR_E_ARG (-1)        /* null pointer / non-positive size / misaligned buffer */
#define CLAIR_E_LIMIT (-2)      /* a CLAIR_MAX_* limit exceeded */
#define CLAIR_E_MODE (-3)       /* unknown interpolation mode or flag combination */

/* interpolation modes, values of clair_torch.common.enums.InterpMode (common/enums.py:10-17) */
#define CLAIR_INTERP_LOOKUP 1
#define CLAIR_INTERP_LINEAR 2
#define CLAIR_INTERP_CATMULL 3

CLAIR_API int clair_abi_version(void);
CLAIR_API const char *clair_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
CLAIR_API uint64_t clair_launch_count(void);

/* developer knob for kernel tuning experiments (keys: hdr_vec, hdr_waves, hdr_force_dynamic, hdr_fixed_max, hdr_prefetch, hdr_tma, fwd_blocks, dark_strip, dark_rows, aux_waves, stats_waves, grad_waves,
 * stats_blocks_per_sm, stats_warps, stats_slots, stats_buffers, grad_blocks_per_sm, grad_pix, grad_warps, grad_copies); 0 restores the
 * library default.  Not part of the reference boundary. */
CLAIR_API int clair_set_tuning(const char *key, int value);

/*
 * ICRF evaluation — replaces ICRFModelBase.forward (models/base.py:135-226): LINEAR (:160-182, rows per
 * Q1), LOOKUP (:138-158, round-half-even, true channel) or CATMULL (:184-226, four taps, rows per Q1).
 * x_dev is (n_frames, C, plane).
 *   y_dev      f(x)                                        (required)
 *   dydx_dev   autograd's d f / d x, LINEAR / CATMULL      (optional, may be NULL)
 */
CLAIR_API int clair_icrf_forward(const float *x_dev, const float *theta_dev, float *y_dev, float *dydx_dev,
                       int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                       const int32_t *curve_row_base_host, void *stream);

/*
 * Back-propagation of clair_icrf_forward to the table: grad_theta[u, x0] += g*(1-w), grad_theta[u, x1] += g*w
 * (the index_put of models/base.py:176 under autograd; four taps at :219 for CATMULL).  grad_theta_dev is (C, L)
 * float64 and is ACCUMULATED into (zero it first).  interp_mode: CLAIR_INTERP_LINEAR, CLAIR_INTERP_CATMULL or
 * CLAIR_INTERP_LOOKUP (one tap with weight 1 in the true channel row: the gather of models/base.py:158).
 */
CLAIR_API int clair_icrf_backward_theta(const float *x_dev, const float *grad_y_dev, double *grad_theta_dev,
                              int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                              const int32_t *curve_row_base_host, void *workspace_dev, size_t workspace_bytes,
                              void *stream);

/* Bytes of device scratch the two table-gradient entry points need (replicated fp32 tables the per-element
 * reductions are spread over); the buffer must be 16-byte aligned and is zeroed by the call itself. */
CLAIR_API size_t clair_grad_workspace_bytes(int n_channels, int lut_size);

/*
 * Linearisation of single images with uncertainty — replaces the core of linearize_dataset_generator
 * (inference/linearization.py:94-106,132): lin = f(x), sigma = sqrt((f'(x) * std)^2); sigma = 0 when
 * std_dev is NULL (:97).  Inputs are (n_frames, C, plane); each frame is an independent image.  interp_mode is the
 * model's InterpMode; CLAIR_INTERP_LOOKUP has no derivative, so std_dev must be NULL with it (the reference's
 * autograd call raises in that combination).
 */
CLAIR_API int clair_linearize(const float *val_dev, const float *std_dev, const float *theta_dev, float *lin_dev,
                    float *sigma_dev, int n_frames, int n_channels, int64_t plane, int lut_size, int interp_mode,
                    const int32_t *curve_row_base_host, void *stream);

/*
 * clair_linearize straight from the camera's integer codes — the reference's CPU transforms CastTo + Normalize
 * (common/transforms.py:107-190: x = fl32(code) / fl32(code_max), an IEEE division) and the missing-std synthesis of
 * MultiFileMapDataset (datasets/base.py:128-133) are evaluated in the kernel's load, so a frame is read as 1 or 2 bytes
 * per sample.  Bit-identical to clair_linearize on the CPU-transformed fp32 image.  LINEAR models only.
 *   codes_dev   (n_frames, C, plane) uint8 (code_bytes = 1) or uint16 (code_bytes = 2); plane a multiple of 4.  May be
 *               page-locked host memory (mapped): the kernel then reads it over PCIe.
 *   std_mode    0 none (sigma = 0), 1 std_dev tensor (n_frames, C, plane) fp32, 2 std = x * std_value, 3 std = std_value
 */
CLAIR_API int clair_linearize_codes(const void *codes_dev, int code_bytes, float code_max, const float *std_dev, int std_mode,
                          float std_value, const float *theta_dev, float *lin_dev, float *sigma_dev, int n_frames,
                          int n_channels, int64_t plane, int lut_size, int interp_mode, const int32_t *curve_row_base_host,
                          void *stream);

/*
 * Camera codes -> the normalised fp32 value stack (and its synthesised std stack) on the device: the reference's CastTo +
 * Normalize CPU transforms (common/transforms.py:107-190: x = fl32(code) / fl32(code_max), an IEEE division) and the
 * missing-std synthesis of MultiFileMapDataset (datasets/base.py:128-133), for the drivers whose kernels take fp32 stacks —
 * measure_linearity and train_icrf (inference/measure_linearity.py:41-43, training/icrf_training.py:96-98): the stack then
 * crosses PCIe as 1 or 2 bytes per sample.  Bit-identical to the CPU transforms.
 *   codes_dev   n_elements uint8 / uint16 codes (any shape, n_elements a multiple of 4), device memory
 *   std_mode    0: no std (std_dev NULL) | 2: std = x * std_value | 3: std = std_value
 */
CLAIR_API int clair_expand_codes(const void *codes_dev, int code_bytes, float code_max, int std_mode, float std_value,
                       int64_t n_elements, float *val_dev, float *std_dev, void *stream);

/*
 * One band of every (frame, channel) slab of a page-locked HOST stack to a dense device buffer, as ONE strided copy
 * (cudaMemcpy2DAsync): the collated batch's .to(device) (inference/measure_linearity.py:41-43) cut into row bands, so that
 * the copy of band b+1 overlaps clair_expand_codes + clair_pair_stats on band b (clair_torch_b200.measure_linearity does
 * that for page-locked code batches).  Byte counts, so any sample type.
 *   dst_dev            n_slabs x band_bytes, dense
 *   src_host           n_slabs slabs of slab_stride_bytes each; the band starts band_offset_bytes into every slab
 */
CLAIR_API int clair_copy_band_h2d(void *dst_dev, const void *src_host, int64_t n_slabs, int64_t slab_stride_bytes,
                                  int64_t band_offset_bytes, int64_t band_bytes, void *stream);

/*
 * clair_linearize for images that live in page-locked HOST memory and whose results are wanted there too (the
 * generator ends with .cpu(), inference/linearization.py:132): the planes are cut into n_bands bands of pixels; band b+1
 * is copied in (in_stream), band b is linearised (stream) and band b-1 is copied out (out_stream) at the same time, so
 * both PCIe directions stay busy.  *_stage_dev are device buffers shaped like the host arrays.  The three streams must
 * differ; when the call returns, `stream` has been made to wait for the last device-to-host copy.
 */
CLAIR_API int clair_linearize_staged(const float *val_host, const float *std_host, float *lin_host, float *sigma_host,
                           float *val_stage_dev, float *std_stage_dev, float *lin_stage_dev, float *sigma_stage_dev,
                           const float *theta_dev, int n_frames, int n_channels, int64_t plane, int lut_size,
                           int interp_mode, const int32_t *curve_row_base_host, int n_bands, void *in_stream,
                           void *out_stream, void *stream);

/*
 * One DataLoader batch of the exposure-weighted HDR merge with first-order uncertainty — replaces the
 * loop body of compute_hdr_image (inference/hdr_merge.py:95-128) including WBOMean.update_values
 * (common/statistics.py:64-109) and the autograd pass at :107-115, in closed form.
 *
 *   val_dev, std_dev      (n_frames, C, plane) fp32, frames sorted by ascending exposure; std_dev may be NULL
 *                         (then no variance is produced)
 *   exposure_host         n_frames exposure times in seconds (float64, the collated 'exposure_time')
 *   theta_dev             (C, lut_size) ICRF table, evaluated in LINEAR mode (other modes: clair_hdr_merge below);
 *                         NULL = identity (icrf_model=None, :99-100)
 *   gaussian_weights      1: w = exp(-30 (x-0.5)^2) (:95 with weight_fn != None); 0: w = 1
 *   mean_state_dev        (C, plane) float64   running weighted mean      } read unless is_first,
 *   wsum_state_dev        (C, plane) float32   running sum of weights     } written unless is_final;
 *   var_state_dev         (C, plane) float32   running variance           } may be NULL if is_first && is_final
 *   is_first / is_final   first / last batch of the stack
 *   radiance_dev          written when is_final: (C, plane), float64 if radiance_f64 else float32
 *                         (the reference returns float64 by type promotion)
 *   sigma_dev             written when is_final and std_dev != NULL: sqrt(variance), (C, plane) float32
 */
CLAIR_API int clair_hdr_merge_update(const float *val_dev, const float *std_dev, const double *exposure_host,
                           int n_frames, const float *theta_dev, int n_channels, int lut_size, int64_t plane,
                           const int32_t *curve_row_base_host, int gaussian_weights,
                           double *mean_state_dev, float *wsum_state_dev, float *var_state_dev,
                           int is_first, int is_final, void *radiance_dev, int radiance_f64,
                           float *sigma_dev, void *stream);

/*
 * Integer-ingest form of clair_hdr_merge_update — SURVEY.md §8(f) rank 2: the reference's CastTo(float32) +
 * Normalize(max_val=code_max, min_val=0) transforms (clair_torch/common/transforms.py:107-190,
 * clair_torch/common/general_functions.py:359-388) and its synthesis of missing std images
 * (clair_torch/datasets/base.py:128-133) are fused into the kernel's load, so the raw camera codes cross PCIe / HBM
 * (1 or 2 bytes per sample instead of 4 + 4).
 *   codes_dev        (n_frames, C, plane) uint8 (code_bytes = 1) or uint16 (code_bytes = 2); H*W must be a multiple of 4
 *   code_max         x = fl32(code) / fl32(code_max), an IEEE fp32 division exactly as the CPU transform computes it
 *   std_mode         0 none | 1 fp32 tensor std_dev | 2 std = x * std_value (MissingStdMode.MULTIPLIER) |
 *                    3 std = std_value (MissingStdMode.CONSTANT)
 * All other arguments as clair_hdr_merge_update.
 */
#define CLAIR_STD_NONE 0
#define CLAIR_STD_TENSOR 1
#define CLAIR_STD_MULTIPLIER 2
#define CLAIR_STD_CONSTANT 3
CLAIR_API int clair_hdr_merge_codes(const void *codes_dev, int code_bytes, float code_max, const float *std_dev, int std_mode,
                          float std_value, const double *exposure_host, int n_frames, const float *theta_dev,
                          int n_channels, int lut_size, int64_t plane, const int32_t *curve_row_base_host,
                          int gaussian_weights, double *mean_state_dev, float *wsum_state_dev, float *var_state_dev,
                          int is_first, int is_final, void *radiance_dev, int radiance_f64, float *sigma_dev,
                          void *stream);

/*
 * Descriptor form of the two entry points above, with the two options they do not carry:
 *   interp_mode    CLAIR_INTERP_LINEAR (fused fast kernels), CLAIR_INTERP_LOOKUP (models/base.py:138-158: nearest sample of
 *                  the true channel row, no derivative: the uncertainty is the weight-derivative term only) or
 *                  CLAIR_INTERP_CATMULL (models/base.py:184-226) — ICRFModelBase.interpolation_mode of the model
 *                  compute_hdr_image was given (inference/hdr_merge.py:99-100)
 *   plane_stride   elements between consecutive channel planes in EVERY buffer (inputs, state, outputs); 0 = plane.
 *                  With plane_stride > plane the call processes `plane` pixels of larger planes (a band of rows of a
 *                  frame): all pointers address the band's first pixel and curve_row_base_host[c] must be
 *                  (c * plane_stride + first_pixel) mod C.
 *   code_bytes     0: val_dev is fp32 and std_dev (may be NULL) its std, std_mode ignored;
 *                  1 / 2: val_dev holds uint8 / uint16 codes, see clair_hdr_merge_codes
 *   dark_dev ...   dark-field correction of the batch fused into the load (inference/hdr_merge.py:76-92,117-126): the
 *                  mixed images and their effective std are formed in registers instead of by a pre-pass
 *   code_layout    CLAIR_CODES_PLANAR: val_dev is (n_frames, C, plane).  CLAIR_CODES_HWC_BGR (integer codes, C = 3, plane a
 *                  multiple of 4): val_dev is (n_frames, H, W, 3) exactly as OpenCV hands a colour image over, and the
 *                  CvToTorch transform (clair_torch/common/general_functions.py:315-336: BGR -> RGB, HWC -> CHW) happens in
 *                  the kernel's address arithmetic; std_dev, state and outputs stay planar RGB.
 * struct_bytes must be sizeof(clair_merge_desc); all other fields as the arguments of clair_hdr_merge_update.
 */
#define CLAIR_CODES_PLANAR 0
#define CLAIR_CODES_HWC_BGR 1
typedef struct clair_merge_desc {
    uint32_t struct_bytes;
    int32_t code_bytes;
    const void *val_dev;
    const float *std_dev;
    int32_t std_mode;
    float std_value;
    float code_max;
    int32_t n_frames;
    const double *exposure_host;
    const float *theta_dev;
    int32_t n_channels;
    int32_t lut_size;
    int32_t interp_mode;
    int32_t gaussian_weights;
    int64_t plane;
    int64_t plane_stride;
    const int32_t *curve_row_base_host;
    double *mean_state_dev;
    float *wsum_state_dev;
    float *var_state_dev;
    int32_t is_first;
    int32_t is_final;
    int32_t radiance_f64;
    int32_t code_layout;        /* CLAIR_CODES_PLANAR, or CLAIR_CODES_HWC_BGR for uint8 / uint16 codes */
    void *radiance_dev;
    float *sigma_dev;
    /* fused dark-field correction (NULL dark_dev = off): see clair_dark_field_mix below for the arithmetic.  dark_dev /
     * dark_std_dev are shaped like val_dev; height * width must equal plane.  Covered: fp32 images with std and dark std,
     * LINEAR model or none, <= 8 frames per batch, even width, dense planes, 16-byte aligned buffers — otherwise the call
     * returns CLAIR_E_MODE and the clair_dark_field_mix pre-pass is the way. */
    const float *dark_dev;
    const float *dark_std_dev;
    int32_t height;
    int32_t width;
    float dark_threshold;
    float dark_alpha;
} clair_merge_desc;

CLAIR_API int clair_hdr_merge(const clair_merge_desc *desc, void *stream);

/*
 * The same merge for a stack that lives in page-locked HOST memory (the collated DataLoader batch of
 * inference/hdr_merge.py:61-66 before its `.to(device)`): val_host / std_host are (n_frames, C, plane) like the device
 * form; desc->val_dev / desc->std_dev name device STAGING buffers of the same size.  The planes are cut into n_bands
 * bands of pixels; the copy engine moves band b+1 (cudaMemcpy2DAsync on copy_stream) while the kernel merges band b on
 * `stream`, so the call costs the PCIe transfer plus one band's kernel.  Outputs and running state (desc->radiance_dev,
 * sigma_dev, *_state_dev) may be device or page-locked host memory.  copy_stream must be a different stream.
 */
CLAIR_API int clair_hdr_merge_staged(const clair_merge_desc *desc, const void *val_host, const float *std_host, int n_bands,
                           void *copy_stream, void *stream);

/*
 * Dark-field correction pre-pass — SURVEY.md §8(f) rank 1.  Replaces conditional_gaussian_blur(images, dark, 0.05, 3,
 * differentiable=True) (clair_torch/common/general_functions.py:440-486) as called at clair_torch/inference/hdr_merge.py:
 * 89-92 and clair_torch/inference/linearization.py:88-91, together with BOTH variance terms the drivers form afterwards
 * (hdr_merge.py:107-126, linearization.py:98-116): the drivers differentiate with respect to the mixed image, so
 *     val_out = m B(x) + (1 - m) x,    m = sigmoid(alpha (dark - threshold)),   B = 3x3 Gaussian blur, reflect padding
 *     std_out = sqrt(std^2 + ((B(x) - x) alpha m (1 - m) dark_std)^2)
 * and running the merge / linearise kernels on (val_out, std_out) reproduces mean, image-std and dark-std terms.
 * All stacks (n_frames, C, H, W) fp32; std_out_dev may be NULL (then std_dev / dark_std_dev are not read).
 */
CLAIR_API int clair_dark_field_mix(const float *val_dev, const float *std_dev, const float *dark_dev, const float *dark_std_dev,
                         int n_frames, int n_channels, int height, int width, float threshold, float alpha,
                         float *val_out_dev, float *std_out_dev, void *stream);

/*
 * Flat-field correction in place — replaces flat_field_mean(F, 1.0) + flatfield_correction
 * (clair_torch/common/general_functions.py:182-238) and the variance term of clair_torch/inference/hdr_merge.py:131-153
 * (mean_in_graph = 1: the whole-image mean of F is differentiated through) or clair_torch/inference/linearization.py:
 * 118-130 (mean_in_graph = 0: the mean is a constant) — SURVEY.md Q11.
 *   value_dev   (n_images, C, plane) fp32 or fp64 (value_f64), overwritten with v / (F + 1e-6) * mean(F)
 *   sigma_dev   (n_images, C, plane) fp32 std, overwritten with sqrt(sigma^2 + (dy/dF flat_std)^2); may be NULL
 *   flat_dev, flat_std_dev   (C, plane) fp32; flat_std_dev may be NULL (value only)
 *   scratch_dev 2*C doubles
 */
CLAIR_API int clair_flat_field_correct(void *value_dev, int value_f64, float *sigma_dev, const float *flat_dev,
                             const float *flat_std_dev, int n_images, int n_channels, int64_t plane, int mean_in_graph,
                             double *scratch_dev, void *stream);

/*
 * Streaming weighted mean / second moment over frames — replaces WBOMeanVar.update_values + _update_internal_values
 * (clair_torch/common/statistics.py:209-259) as used by compute_video_mean_and_std
 * (clair_torch/inference/inferential_statistics.py:19-49); SURVEY.md §8(f) rank 3.
 *   val_dev       (n_frames, C, plane) fp32 batch of frames (any n_frames)
 *   weights_dev   same shape, or NULL (then W_B = n_frames, as statistics.py:226-229)
 *   theta_dev     (C, L) table to linearise the frames first (interp_mode = the model's InterpMode), or NULL
 *   mean / m2 / wsum / wsq state   (C, plane) fp32 each: running mean, M2, sum of weights, sum of squared weights;
 *                 overwritten when is_first, merged (Chan / West) otherwise
 */
CLAIR_API int clair_frame_stats_update(const float *val_dev, const float *weights_dev, const float *theta_dev, int n_frames,
                             int n_channels, int64_t plane, int lut_size, int interp_mode,
                             const int32_t *curve_row_base_host,
                             float *mean_state_dev, float *m2_state_dev, float *wsum_state_dev, float *wsq_state_dev,
                             int is_first, void *stream);

/*
 * Pairwise exposure-ratio statistics — replaces, for one batch, get_pairwise_valid_pixel_mask
 * (common/general_functions.py:276-312), combined_gaussian_pair_weights (training/losses.py:208-235),
 * pixelwise_linearity_loss (:13-67) and the reductions of compute_spatial_linearity_loss (:70-108) /
 * weighted_mean_and_std (common/general_functions.py:118-178), as used by measure_linearity
 * (inference/measure_linearity.py:41-74) and the forward half of a train_icrf step
 * (training/icrf_training.py:105-136).
 *
 *   pair_i_host, pair_j_host, pair_ratio_host   P pairs from get_valid_exposure_pairs (int32, int32, float64)
 *   theta_dev            (C, L) table or NULL (identity linearisation, measure_linearity.py:53)
 *   interp_mode          the model's InterpMode (ignored when theta_dev is NULL): CLAIR_INTERP_LINEAR (models/base.py:160-182),
 *                        CLAIR_INTERP_CATMULL (:184-226) or CLAIR_INTERP_LOOKUP (:138-158).  A LOOKUP model has no derivative
 *                        with respect to the image, so std_dev must be NULL with it (CLAIR_E_MODE; the reference's
 *                        autograd.grad raises at measure_linearity.py:57-63 / icrf_training.py:117-124)
 *   valid_lo, valid_hi   inclusive validity range, compared in fp32 (training default 1/255, 254/255)
 *   relative             use_relative_linearity_loss
 *   unc_weighting        use_uncertainty_weighting (only has an effect when std_dev != NULL)
 *   sums_dev             (P, C, 5) float64, ACCUMULATED into (zero it first; shards of one image may add
 *                        into the same buffer, or be all-reduced):
 *                          [0] sum M*Wt   [1] sum M*Wt*l   [2] sum M*Wt*l^2   [3] sum M*err   [4] sum M
 *                        from which  mean = s1/max(s0,1e-8),  std = sqrt(max(s2 - 2 mean s1 + mean^2 s0, 0)/max(s0,1e-8)),
 *                        errmean = s3/max(s4,1e-8).
 */
CLAIR_API int clair_pair_stats(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                     const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                     int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                     const int32_t *curve_row_base_host,
                     float valid_lo, float valid_hi, int relative, int unc_weighting, double *sums_dev,
                     void *stream);

/*
 * Training-step variant of clair_pair_stats: same arguments, but only sums [0] (sum M*Wt) and [1] (sum M*Wt*l) are
 * produced — all a train_icrf step needs (training/icrf_training.py:133-136) — and the std images are only read when
 * the uncertainty weights need them.
 */
CLAIR_API int clair_pair_means(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                     const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                     int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                     const int32_t *curve_row_base_host,
                     float valid_lo, float valid_hi, int relative, int unc_weighting, double *sums_dev,
                     void *stream);

/*
 * (P, C)-sized algebra between the two passes of a training step (training/icrf_training.py:133-136 and the head of
 * the closed-form backward), on the device so the step never synchronises:
 *   mean[p,c] = s1/max(s0,1e-8);  linloss[c] = sqrt(sum_p mean[p,c]^2);
 *   upstream[p,c] = mean/linloss/max(s0,1e-8) (0 where linloss = 0);  mean_for_grad = mean (0 where s0 < 1e-8).
 * All buffers float64: sums (P,C,5) in; linloss (C), mean / upstream / mean_for_grad (P,C) out.
 */
CLAIR_API int clair_pair_upstream(const double *sums_dev, int n_pairs, int n_channels, double *linloss_dev, double *mean_dev,
                        double *upstream_dev, double *mean_for_grad_dev, void *stream);

/*
 * The four curve penalties of training/losses.py:111-190 per channel, weighted as in training/icrf_training.py:143,
 *   penalty[c] = alpha*monotonicity + beta*range + gamma*endpoints + delta*smoothness     (float64 (C), overwritten)
 * and their gradient ADDED to grad_theta_dev (C, L) float64.  fp32 arithmetic like the reference.
 */
CLAIR_API int clair_curve_penalties(const float *theta_dev, int n_channels, int lut_size, float alpha, float beta, float gamma,
                          float delta, double *penalty_dev, double *grad_theta_dev, void *stream);

/*
 * Gradient of the linearity loss of one train_icrf step with respect to the ICRF table — replaces the C
 * `loss[c].backward(retain_graph=True)` passes (training/icrf_training.py:148-149) for the linearity term,
 * in closed form.  With CLAIR_INTERP_CATMULL the upstream of each frame element goes to its four
 * Catmull-Rom taps (models/base.py:219-226), with CLAIR_INTERP_LOOKUP to its nearest sample in the true channel row
 * (:158).  Arguments as clair_pair_stats, plus
 *   upstream_dev    (P, C) float64: U[p,c] = dLoss_c/dmean[p,c] / max(s0[p,c],1e-8)
 *   mean_dev        (P, C) float64: the spatial means (needed when the weights depend on the curve)
 *   grad_theta_dev  (C, L) float64, ACCUMULATED into (zero it first).
 *   workspace_dev   clair_grad_workspace_bytes(C, L) bytes of scratch
 */
CLAIR_API int clair_pair_grad(const float *val_dev, const float *std_dev, int n_frames, int n_channels, int64_t plane,
                    const int32_t *pair_i_host, const int32_t *pair_j_host, const double *pair_ratio_host,
                    int n_pairs, const float *theta_dev, int lut_size, int interp_mode,
                    const int32_t *curve_row_base_host,
                    float valid_lo, float valid_hi, int relative, int unc_weighting,
                    const double *upstream_dev, const double *mean_dev, double *grad_theta_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);

/*
 * One training step's statistics AND table gradient in a single pass over the stack, for the case of exactly one exposure
 * pair without uncertainty weighting (BASELINE config 5: data-parallel training on 100 MP exposure pairs; replaces
 * clair_pair_means + clair_pair_upstream + clair_pair_grad for training/icrf_training.py:105-149).  With one pair the
 * upstream factor of the closed-form backward is one scalar per channel, so the kernel scatters the un-normalised gradient
 * and the scalar is applied afterwards:
 *   fused_dev   clair_pair_fused_doubles(C, L) float64, ACCUMULATED into (zero it first):
 *               [c*5 + 0] sum M*Wt, [c*5 + 1] sum M*Wt*l of channel c (the (1, C, 5) layout of clair_pair_stats), then
 *               T[c][u][k] (C, C, L): the un-normalised table gradient of the elements of channel c.
 *               Row-band shards of one image add (all-reduce) this ONE buffer.
 *   clair_pair_fused_combine: mean[c] = s1/max(s0,1e-8), linloss[c] = sqrt(mean^2), U[c] = mean/linloss/max(s0,1e-8),
 *               grad_theta[u][k] += sum_c U[c] T[c][u][k]   (grad_theta_dev (C, L) float64, may be NULL: loss only).
 * Other arguments as clair_pair_grad.  n_pairs must be 1 (CLAIR_E_MODE otherwise); H*W even, 8-byte aligned stack.
 */
CLAIR_API size_t clair_pair_fused_doubles(int n_channels, int lut_size);
CLAIR_API int clair_pair_fused(const float *val_dev, int n_frames, int n_channels, int64_t plane, const int32_t *pair_i_host,
                     const int32_t *pair_j_host, const double *pair_ratio_host, int n_pairs, const float *theta_dev,
                     int lut_size, int interp_mode, const int32_t *curve_row_base_host, float valid_lo, float valid_hi,
                     int relative, double *fused_dev, void *workspace_dev, size_t workspace_bytes, void *stream);
CLAIR_API int clair_pair_fused_combine(const double *fused_dev, int n_channels, int lut_size, double *linloss_dev,
                             double *mean_dev, double *grad_theta_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CLAIR_B200_H */
