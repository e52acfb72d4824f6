/*
 * mriacl_recon.h -- C ABI of libmriacl_recon.so, the B200 (sm_100a) k-space -> image
 * input stage for bonhchi/mri_acl_imagesegmentation_adsp.
 *
 * The reference has no FFI layer: its operator API for this path is a set of plain
 * Python functions.  Each entry point below names the
 * reference function(s) it replaces; REF = the reference checkout, ZIP! = a member of
 * REF/reference/fastMRI_prostate-main.zip.  The Python side
 * (mri_acl_imagesegmentation_adsp_b200/adapters/recon_cabi.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types cross the boundary.
 *  - every data pointer is a DEVICE pointer on the current CUDA device unless the
 *    parameter says HOST; the caller owns all buffers, including the workspace.
 *  - complex64 = interleaved (re, im) float32, i.e. numpy complex64 / torch complex64 /
 *    the fastMRI real view (..., 2) -- all byte-identical.
 *  - all work is enqueued on `cuda_stream` (a cudaStream_t; NULL = legacy default
 *    stream) and returns without synchronising.  Sampling masks are HOST vectors: the
 *    library turns each distinct mask into a device-resident plan once and caches it.
 *  - return value: MRIACL_OK (0) or a negative MRIACL_ERR_* code; the message is
 *    available from mriacl_last_error() (thread-local).  No C++ exception escapes.
 *  - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef MRIACL_RECON_H
#define MRIACL_RECON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRIACL_ABI_VERSION 1

/* return codes */
#define MRIACL_OK               0
#define MRIACL_ERR_INVALID     -1   /* bad shape / argument  -> Python ValueError  */
#define MRIACL_ERR_UNSUPPORTED -2   /* size outside plan limits -> ValueError       */
#define MRIACL_ERR_CUDA        -3   /* CUDA runtime failure  -> RuntimeError        */
#define MRIACL_ERR_WORKSPACE   -4   /* workspace too small   -> RuntimeError        */

/* flags for mriacl_recon_rss_f32 */
#define MRIACL_FLIP_ROWS      0x1u  /* np.flipud of each combined image (ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:101) */
#define MRIACL_NORM_INSTANCE  0x2u  /* (x-mean)/(std+eps), unbiased std (ZIP!/DL_reconstruction/data/transforms.py:143-162) */
#define MRIACL_FORCE_GENERIC  0x4u  /* testing: use the generic (any-size) kernels even where a fused plan exists */
/* kernel schedule of the fused 640x368 plan.  The product library holds ONE schedule, MRIACL_SEQUENTIAL (the default);
 * the MRIACL_SCHED_* schedules below were measured slower, exist only in builds with
 * -DMRIACL_EXPERIMENTAL (libmriacl_recon_exp.so) and return MRIACL_ERR_UNSUPPORTED from the product library. */
#define MRIACL_SEQUENTIAL     0x8u   /* column pass -> row pass -> normalise, back to back on the caller's stream */
#define MRIACL_SCHED_FUSED    0x10u  /* experimental: one persistent kernel, CTAs switch between column and row items */
#define MRIACL_SCHED_CORESIDENT 0x80u /* one persistent kernel, one CTA per SM: column team + row team co-resident, normalisation fused */
#define MRIACL_SCHED_PIPELINED 0x800u /* sequential kernels on small chunks, two streams, T double-buffered in L2 (eviction hints) */
#define MRIACL_SCHED_PAIR     0x40u  /* column pass -> pair row pass (rowpair.cuh: per-output-pair, barrier-free) -> normalise */
#define MRIACL_SCHED_OVERLAP  0x20u  /* experimental: persistent row pass on a side stream fed by per-slice counters */
/* profiling only (bench.py's per-kernel timing): run just the named phase(s) of the fused plan;
 * none set = the whole stage.  The workspace must still hold the previous phase's output. */
#define MRIACL_ONLY_COLPASS   0x100u
#define MRIACL_ONLY_ROWPASS   0x200u
#define MRIACL_ONLY_NORM      0x400u
/* kspace holds only the SAMPLED columns (mask value != 0), densely: element (b,a,c,h,j) of sampled column j lives at
 * b*slice_stride + a*avg_stride + (c*H + h)*n_act + j, as written by mriacl_pack_columns_host.  W, pad_left and mask
 * still describe the full line.  Plans with the 640-row column pass only (MRIACL_ERR_UNSUPPORTED otherwise). */
#define MRIACL_PACKED_COLUMNS 0x1000u

/* path selector reported by mriacl_supported */
#define MRIACL_PATH_NONE    0
#define MRIACL_PATH_GENERIC 1
#define MRIACL_PATH_FUSED   2

int         mriacl_abi_version(void);
const char* mriacl_last_error(void);

/* Which kernels serve a (H, W_padded) transform: MRIACL_PATH_FUSED for the shapes with a
 * hand-scheduled fused plan (640x368 knee), MRIACL_PATH_GENERIC
 * for any other size up to MRIACL_MAX_LINE per axis, MRIACL_PATH_NONE beyond that. */
#define MRIACL_MAX_LINE 4096
int mriacl_supported(int H, int W_padded);

/* Number of kernels this library has launched in the calling process (all threads).
 * bench.py reports the delta over its timed region as "gpu_launches". */
uint64_t mriacl_launch_count(void);

/*
 * The fused stage:  mask apply -> zero-pad PE -> centred 2-D iFFT per coil -> RSS over
 * coils -> [flipud] -> mean over averages -> centre crop -> [instance normalise].
 *
 * Replaces, in one call per batch, the chain the reference spells three ways:
 *   REF/src/utils/kspace.py:11-31 (ifft2c, complex_abs, center_crop_or_pad) + sqrt(sum^2);
 *   ZIP!/DL_reconstruction/fftc.py:41-65 (ifft2c_new), coil_combine.py:28-41 (rss_complex),
 *        data/transforms.py:45-67 (center_crop), :143-162 (normalize_instance);
 *   ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-75,80-121 and
 *        fastmri_prostate/data/mri_data.py:123-160 (zero_pad_kspace_hdr).
 *
 * kspace   complex64, logical shape [B, A, C, H, W]; element (b,a,c,h,w) lives at
 *          b*slice_stride + a*avg_stride + (c*H + h)*W + w   (strides in complex elements;
 *          knee batches: A=1, slice_stride=C*H*W; prostate files (A,S,C,RO,PE):
 *          slice_stride=C*H*W, avg_stride=S*C*H*W).
 * mask_w   HOST float32[W] multiplied along the last axis, or NULL for all ones.
 *          Columns whose mask value is exactly 0 are skipped (treated as exact zeros).
 * pad_left, W_padded   zero-pad the PE axis to W_padded with pad_left zeros on the left
 *          (W_padded == W and pad_left == 0 for no padding).
 * out      float32 [B, out_h, out_w]; out_h <= H, out_w <= W_padded (centre crop,
 *          start (n-out)/2).  mean_std: float32 [B, 2] (mean, unbiased std of the
 *          un-normalised crop) or NULL.
 * workspace  device scratch of at least mriacl_recon_rss_workspace_bytes(.., slices=1)
 *          bytes, 256-byte aligned; a larger one lets more slices be in flight per launch.
 */
size_t mriacl_recon_rss_workspace_bytes(int slices, int A, int C, int H, int W, int pad_left, int W_padded,
                                        int out_h, int out_w, const float* mask_w_host, unsigned flags);

int mriacl_recon_rss_f32(const void* kspace_c64, long long slice_stride, long long avg_stride,
                         const float* mask_w_host, float* out, float* mean_std,
                         int B, int A, int C, int H, int W, int pad_left, int W_padded,
                         int out_h, int out_w, unsigned flags, float eps,
                         void* workspace, size_t workspace_bytes, void* cuda_stream);

/* HOST helper of the end-to-end path: copy the sampled columns (mask value != 0; NULL mask = all) of n_rows rows of
 * W complex64 values from kspace_c64_host to packed_c64_host [n_rows][n_act], on n_threads host threads (<= 0: one per
 * core this process may run on).  Unsampled columns are multiplied by zero by the mask apply (fastMRI convention,
 * SURVEY.md section 0 fact 4), so they need not cross PCIe: the caller copies the packed buffer to the device and
 * passes MRIACL_PACKED_COLUMNS.  Both pointers are HOST pointers (pageable or pinned); no CUDA call is made.
 * Returns n_act (>= 0) or a negative MRIACL_ERR_* code. */
int mriacl_pack_columns_host(const void* kspace_c64_host, void* packed_c64_host, long long n_rows, int W,
                             const float* mask_w_host, int n_threads);

/* |ifft2c(k)| as float32 for B single-coil (H, W) slices.
 * Replaces MRIKneePreprocessor.ifft2c_single, REF/src/preprocess/mri_preprocess.py:149-160.
 * workspace: B*H*W*8 bytes (complex scratch) unless (H, W) has a fused plan, where
 * mriacl_recon_rss_workspace_bytes(B,1,1,H,W,0,W,H,W,NULL,0) applies. */
size_t mriacl_ifft2c_abs_workspace_bytes(int B, int H, int W);
int mriacl_ifft2c_abs_f32(const void* kspace_c64, float* out, int B, int H, int W,
                          void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Centred orthonormal 2-D FFT (inverse != 0: iFFT) over the last two axes of [B, H, W]
 * complex64, out of place or in place (in == out).
 * Replaces fft2c / ifft2c, REF/src/utils/kspace.py:4-16; fft2c_new / ifft2c_new,
 * ZIP!/DL_reconstruction/fftc.py:14-65; ifftnd over two axes,
 * ZIP!/fastmri_prostate/reconstruction/utils.py:7-29. */
int mriacl_fft2c_c64(const void* in_c64, void* out_c64, int B, int H, int W, int inverse, void* cuda_stream);

/* |z| (squared == 0) or |z|^2 (squared != 0) of n complex64 values.
 * Replaces complex_abs, REF/src/utils/kspace.py:18-20; complex_abs / complex_abs_sq,
 * ZIP!/DL_reconstruction/math_fn.py:55-86. */
int mriacl_complex_abs_f32(const void* in_c64, float* out, size_t n, int squared, void* cuda_stream);

/* Root-sum-of-squares over the middle axis of [outer, C, inner]:
 * is_complex == 0: float32 input (rss, ZIP!/DL_reconstruction/coil_combine.py:12-25);
 * is_complex != 0: complex64 input (rss_complex :28-41; rss, prostate_t2_recon.py:105-121). */
int mriacl_rss_f32(const void* in, float* out, size_t outer, int C, size_t inner, int is_complex, void* cuda_stream);

/* Centre crop or zero-pad the last two axes of [B, H, W] to [B, out_h, out_w];
 * elem_bytes is 4 (float32) or 8 (complex64).
 * Replaces center_crop_or_pad, REF/src/utils/kspace.py:22-31 (and, for out <= in,
 * center_crop transforms.py:45-67 / center_crop_im utils.py:54-73). */
int mriacl_center_crop_or_pad(const void* in, void* out, int B, int H, int W, int out_h, int out_w,
                              int elem_bytes, void* cuda_stream);

/* Per-image instance normalisation of [B, n] float32: out = (in-mean)/(std+eps) with the
 * UNBIASED std; mean_std [B,2] or NULL; in == out allowed.
 * Replaces normalize_instance, ZIP!/DL_reconstruction/data/transforms.py:143-162. */
int mriacl_normalize_instance_f32(const float* in, float* out, float* mean_std, int B, size_t n, float eps,
                                  void* cuda_stream);

/* ---- the per-slice steps that follow the reconstruction on the reference's live call path
 * (MRIKneePreprocessor.preprocess_record, REF/src/preprocess/mri_preprocess.py:59-84).  The Otsu / morphology body mask
 * between clip and resize (:194-214) needs scikit-image and stays with the caller: masks are inputs here. ---- */

/* lo = np.percentile(img, pmin), hi = np.percentile(img, pmax) per image of [B, n] float32 (method "linear" in
 * numpy >= 2.0's float32 arithmetic, order statistics selected exactly), out = np.clip(img, lo, hi).
 * out [B, n] or NULL; lo_hi [B, 2] or NULL.  Replaces _percentile_clip, mri_preprocess.py:182-185. */
int mriacl_percentile_clip_f32(const float* in, float* out, float* lo_hi, int B, size_t n, float pmin, float pmax,
                               void* cuda_stream);

/* F.interpolate(size=(out_h, out_w), mode="bilinear", align_corners=False) of [B, H, W] float32.
 * Replaces _resize_np, mri_preprocess.py:187-191. */
int mriacl_resize_bilinear_f32(const float* in, float* out, int B, int H, int W, int out_h, int out_w, void* cuda_stream);

/* the same interpolation of a uint8 mask followed by "> 0.5" (mri_preprocess.py:77): uint8 [B, H, W] -> [B, out_h, out_w]. */
int mriacl_resize_mask_u8(const uint8_t* in, uint8_t* out, int B, int H, int W, int out_h, int out_w, void* cuda_stream);

/* In-mask z-score and [0,1] preview of [B, n] float32 (in may alias out_z).  mask uint8 [B, n] or NULL (all inside).
 * out_z = (x - mean) / std with the mean and POPULATION std of the pixels inside the mask (of the whole image when
 * fewer than 10 are inside), std <= 1e-6 -> 1; out_01 = (x - lo) / float32(hi - lo + 1e-6) with the extrema inside the
 * mask (whole image when it is empty).  stats [B, 6] = mean, std, lo, hi, pixels inside, 1 if the mask was used; any
 * output may be NULL.  Replaces _zscore_in_mask / _preview_01, mri_preprocess.py:216-233. */
int mriacl_zscore_preview_f32(const float* in, const uint8_t* mask, float* out_z, float* out_01, float* stats, int B,
                              size_t n, void* cuda_stream);

/* The three steps as one call on a batch (optional epilogue of the fused stage): percentile clip at full resolution,
 * bilinear resize of the clipped image (and of the body mask, thresholded at 0.5), in-mask z-score + preview.
 * img [B, H, W]; body_mask uint8 [B, H, W] or NULL; out_z [B, out_h, out_w] (required); out_01 or NULL; out_mask uint8
 * (required with body_mask); clip_lo_hi [B, 2] (required, receives the percentiles); stats [B, 6] or NULL. */
int mriacl_clip_resize_zscore_f32(const float* img, const uint8_t* body_mask, float* out_z, float* out_01,
                                  uint8_t* out_mask, float* clip_lo_hi, float* stats, int B, int H, int W, int out_h,
                                  int out_w, float pmin, float pmax, void* cuda_stream);

/* The network-input epilogue of the consumer (KneeNPZ2DSlices.__getitem__, REF/src/dataio/datasets.py:90-95,128-131) for a
 * whole volume: out[s, d] = in[clamp(s + d - k/2, 0, S-1)] (2.5-D stack of k neighbouring slices, k odd), or with repeat != 0
 * out[s, d] = in[s] (one channel repeated for ImageNet encoders); mean_dev / std_dev: DEVICE float32 [k] per output channel,
 * applied as (x - mean) / std, or both NULL.  in [S, n], out [S, k, n]. */
int mriacl_stack25d_f32(const float* in, float* out, int S, size_t n, int k, int repeat, const float* mean_dev,
                        const float* std_dev, void* cuda_stream);

/* GRAPPA weight application, in place, for n_slices slices that share one kernel-geometry plan (one Grappa object of
 * the reference) and carry their own weights.  Replaces Grappa.apply_weights,
 * ZIP!/fastmri_prostate/reconstruction/grappa.py:173-222 (called per (average, slice) by prostate_t2_recon.py:52-63 and
 * dwi/prostate_dwi_recon.py:92-97).  For every hole (x, y) of geometry g: k[x, y, :] += W_g @ S, S = the sampled
 * neighbours of the kx x ky window in window-position-major, coil-minor order.
 *   kspace_inout  complex64, element (slice, x, y, c) at slice*slice_stride + x*sx + y*sy + c*sc (complex elements):
 *                 any axis order of the file works without a transpose.  Only holes are written; the sources of a hole are
 *                 sampled positions, which are never written, so in place is race-free.
 *   The plan tables are DEVICE int32 / int64 arrays built by the host mirror (prostate/grappa.py):
 *   hole_xy [n_holes] = x*Y + y grouped by geometry; item_* [n_items]: geometry, first hole and hole count (<= 256) of each
 *   work item; geom_src_start [n_geom+1] into src_off; src_off = (di + kx/2)*8 + (dj + ky/2) of every sampled window
 *   position; max_sources = largest sources-per-geometry; geom_w_start [n_geom] = offset of W_g (C x n_s*C, row-major)
 *   inside one slice's weight block; weights_c64 [n_slices][weights_per_slice].   C <= 16, kernel at most 7 x 7. */
int mriacl_grappa_apply_c64(void* kspace_inout, long long slice_stride, long long sx, long long sy, long long sc,
                            int n_slices, int X, int Y, int C, int kx, int ky,
                            const int* hole_xy, int n_items, const int* item_geom, const int* item_first,
                            const int* item_count, const int* geom_src_start, const int* src_off, int max_sources,
                            const long long* geom_w_start, const void* weights_c64, long long weights_per_slice,
                            void* cuda_stream);

/* SENSE-style coil combine of [B, C, n] complex64 images with sensitivity maps [B, C, n] (shared_sens != 0: [1, C, n]):
 * sum_c img_c * conj(sens_c); magnitude != 0: float32 |.| [B, n] (np.abs(np.sum(img * sens.conj(), axis=1)),
 * ZIP!/fastmri_prostate/reconstruction/dwi/prostate_dwi_recon.py:106-109), else the complex sum [B, n]
 * (sens_reduce, ZIP!/DL_reconstruction/models/varnet.py:199-203). */
int mriacl_sense_combine(const void* img_c64, const void* sens_c64, void* out, int B, int C, size_t n, int shared_sens,
                         int magnitude, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MRIACL_RECON_H */
