/*
 * tise_b200.h -- C ABI of libtise_b200.so: the B200 (sm_100a) implementation of TISE's
 * image-realism hot path (Inception-v3 features -> FID / O-FID moments + Frechet distance,
 * IS / IS* / O-IS temperature-softmax + KL partial sums).
 *
 * The reference (VinAIResearch/tise-toolbox) is pure Python and has no FFI; the boundary it
 * offers is the Python call surface of image_realism/FID/fid_score.py, inception.py,
 * img_data.py, object_fidelity/O-IS/object_centric_inception_score.py and the two IS*
 * scripts.  Each entry point below names the reference code it replaces (file:line, relative
 * to the reference root); tise_toolbox_b200/*.py binds them with ctypes and re-exposes the
 * reference's Python names.  INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (TISE_OK) or a negative tise_status; nothing throws or exits;
 *     tise_last_error(ctx) holds a human-readable message for the last failure on that ctx.
 *   - "d_" pointers are device pointers on ctx's device, "h_" pointers are host pointers.
 *   - the caller owns every buffer, including workspaces (sizes from *_workspace_bytes);
 *     only the packed network weights are owned by the ctx.
 *   - all work is enqueued on the given stream; no call synchronises the device unless its
 *     comment says so.  One ctx per (device, host thread); there is no global mutable state.
 *   - there is no CPU fallback: without a CUDA device every compute call returns TISE_ERR_CUDA.
 */
#ifndef TISE_B200_H
#define TISE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tise_ctx tise_ctx;
typedef void* tise_stream; /* cudaStream_t */

typedef enum {
  TISE_OK = 0,
  TISE_ERR_INVALID = -1,   /* bad argument (shape, null pointer, unknown weight name) */
  TISE_ERR_CUDA = -2,      /* CUDA runtime / driver error, or no device */
  TISE_ERR_WORKSPACE = -3, /* workspace too small */
  TISE_ERR_STATE = -4,     /* call order (e.g. forward before weights are loaded) */
  TISE_ERR_NUMERIC = -5    /* non-finite intermediate in tise_frechet */
} tise_status;

/* Feature taps, image_realism/FID/inception.py:14-19 (BLOCK_INDEX_BY_DIM). */
typedef enum { TISE_TAP_64 = 0, TISE_TAP_192 = 1, TISE_TAP_768 = 2, TISE_TAP_2048 = 3 } tise_tap;

/* Convolution engine.  TISE_CONV_UMMA is the product path (tcgen05 implicit GEMM);
 * TISE_CONV_SIMT is a slow same-numerics CUDA-core kernel kept for on-device cross checks. */
typedef enum { TISE_CONV_UMMA = 0, TISE_CONV_SIMT = 1 } tise_conv_impl;

int tise_version(void);
const char* tise_status_string(int status);

/* ---- context -------------------------------------------------------------------------- */
int tise_ctx_create(int device, tise_ctx** out);
void tise_ctx_destroy(tise_ctx* ctx);
const char* tise_last_error(const tise_ctx* ctx);

/* conv engine (default TISE_CONV_UMMA) and operand passes: 2 = fp16 activations x (hi+lo) fp16
 * weights, fp32 TMEM accumulation (default; FID within 1e-4 of fp32), 1 = single fp16 pass. */
int tise_set_conv_mode(tise_ctx* ctx, int impl, int passes);
/* mixed operand policy: with passes = 2, convolutions whose INPUT map is at most `single_pass_max_hw`
 * pixels high (0 = none, the default) skip the W_lo pass.  Measured trade-off: DESIGN.md section 2. */
int tise_set_pass_policy(tise_ctx* ctx, int single_pass_max_hw);
/* One-off calibration of the single-pass layers (after tise_load_inception and tise_set_input_lut; synchronous): a strict
 * forward pass over `n_images` synthetic images generated on the device measures the mean activation of every
 * convolution input, and  shift += scale * (W_lo . mean)  becomes the shift of the layer's single-pass form -- the
 * image-independent part of the dropped W_lo product.  `d_ws`: tise_inception_workspace_bytes(ctx, n_images) bytes. */
int tise_calibrate_single_pass(tise_ctx* ctx, int n_images, void* d_ws, size_t ws_bytes, tise_stream stream);
/* Progress marker for callers that overlap host -> device copies with the trunk: every tise_inception_features call
 * records `cuda_event` (a cudaEvent_t; NULL clears) on its stream when the forward pass reaches the first convolution
 * whose input map is at most `max_hw` pixels high.  With max_hw = 17 that is the end of the HBM-bound front of the
 * network (stem ... Mixed_6a): a copy stream that waits for the event moves its traffic into the compute-bound part,
 * where it does not compete for HBM with the layers that run at the memory limit. */
int tise_set_progress_event(tise_ctx* ctx, void* cuda_event, int max_hw);

/* ---- weights -------------------------------------------------------------------------- */
/* One tensor of a torchvision Inception3 state_dict (fp32, host, contiguous), e.g.
 * "Mixed_5b.branch1x1.conv.weight", "....bn.running_var", "fc.weight", "fc.bias".
 * Replaces models.inception_v3(pretrained=True) + load_state_dict at
 * image_realism/FID/inception.py:57 and object_centric_inception_score.py:42-46. */
typedef struct {
  const char* name;
  const float* data;
  int64_t numel;
} tise_weight_desc;

/* Folds eval-mode BatchNorm (eps = 1e-3) into per-channel scale/shift, splits every conv weight
 * into fp16 hi + lo planes in the K-major layout the UMMA kernels read, uploads everything to
 * ctx-owned device memory.  num_classes = rows of fc.weight (0: no classifier head).
 * Synchronises the device. */
int tise_load_inception(tise_ctx* ctx, const tise_weight_desc* tensors, int n_tensors, int num_classes);

/* ---- A1: PNG files -> RGB8 in HBM --------------------------------------------------------- */
/* Replaces PIL.Image.open(f).convert("RGB") on the DataLoader workers
 * (image_realism/FID/img_data.py:19-25, object_centric_inception_score.py:100-102) for PNG files:
 * zlib inflate, scanline unfiltering and conversion to RGB run on the device, one warp per image,
 * and the result is the packed uint8 HWC layout tise_resize_u8 / tise_crop_resize_u8 read.
 * Handled (plain or Adam7-interlaced): bit depth 8 for colour types 0 (grey), 2 (RGB), 3 (palette), 4 (grey+alpha),
 * 6 (RGBA); bit depth 16 for types 2, 4, 6 (the high byte, as Pillow keeps it); bit depths 1 / 2 / 4 for grey
 * (scaled to 0..255) and palette.  Alpha / tRNS are dropped exactly as convert("RGB") drops them.  16-bit grey
 * ("I;16" in Pillow) is reported with supported = 0 and stays with the caller. */
typedef struct {
  int32_t width, height;
  int32_t bit_depth, color_type, interlace;
  int32_t channels;   /* filter unit: bytes per pixel of the scanlines (at least 1) */
  int32_t n_idat;     /* IDAT chunks */
  int32_t supported;  /* 1: tise_png_decode handles this file */
  int64_t idat_bytes; /* compressed bytes in all IDAT chunks */
} tise_png_info;
/* Host only (no ctx, no device): parses the chunk list of one file.  TISE_ERR_INVALID if the bytes are
 * not a well-formed PNG container. */
int tise_png_probe(const uint8_t* h_file, int64_t file_bytes, tise_png_info* info);
size_t tise_png_decode_workspace_bytes(const tise_png_info* infos, int n);
/* n files packed in one buffer that exists both on the host (h_files, only chunk headers are read) and
 * on the device (d_files, same bytes); file i = [h_file_off[i], + h_file_bytes[i]).  Image i is written as
 * height x width x 3 uint8 at d_out + h_out_off[i].  d_status[i] (device int32): 0 ok, > 0 a corrupt
 * stream (bad Huffman code, size mismatch, bad filter byte, Adler-32 mismatch ...), -1 not reached; the
 * caller reads it back after the stream is synchronised.  Chunk CRCs are not verified.
 * TISE_ERR_INVALID if a file is unsupported (see tise_png_probe). */
int tise_png_decode(tise_ctx* ctx, const uint8_t* h_files, const uint8_t* d_files, const int64_t* h_file_off,
                    const int64_t* h_file_bytes, int n, uint8_t* d_out, const int64_t* h_out_off,
                    int32_t* d_status, void* d_ws, size_t ws_bytes, tise_stream stream);

/* Baseline JPEG files, same contract as the PNG entry points above: bit-identical to Pillow's
 * Image.open(f).convert("RGB"), i.e. to libjpeg-turbo with its defaults (jpeg_idct_islow, fancy h2v1 / h2v2
 * upsampling, 16-bit fixed-point YCbCr -> RGB).  Handled: 8-bit Huffman files, sequential (SOF0 / SOF1) or progressive
 * (SOF2), grey or YCbCr, chroma 4:4:4 / 4:2:2 / 4:2:0, restart intervals.  Arithmetic-coded, CMYK and
 * RGB-tagged files are reported with supported = 0 and stay with the caller.  d_status[i]: 0 ok, 21..24 a corrupt
 * entropy-coded segment (csrc/jpeg_core.h), -1 not reached. */
typedef struct {
  int32_t width, height;
  int32_t components;       /* 1 (grey) or 3 (YCbCr) */
  int32_t chroma_mode;      /* 0 grey, 1 4:4:4, 2 4:2:2 (h2v1), 3 4:2:0 (h2v2) */
  int32_t restart_interval; /* MCUs between RSTn markers in the first scan, 0 = none */
  int32_t supported;        /* 1: tise_jpeg_decode handles this file */
  int32_t progressive;      /* SOF2 */
  int32_t scans;            /* SOS segments */
  int32_t huffman_tables;   /* DHT definitions in the file */
  int32_t pad;
  int64_t scan_bytes;       /* bytes from the first entropy-coded byte to the end of the file */
  int64_t blocks;           /* 8 x 8 blocks of all components (whole MCUs) */
} tise_jpeg_info;
int tise_jpeg_probe(const uint8_t* h_file, int64_t file_bytes, tise_jpeg_info* info);

/* ---- host-side ingest helpers (no device code; callable without a GPU) --------------------
 * What img_data.py:19-35's Dataset + DataLoader workers do per file from Python, done per decode GROUP in one call each,
 * on `n_threads` host threads, so that a foreign-function caller holds no interpreter lock meanwhile:
 *   tise_file_sizes  stat() of every path (sizes[i] = -1 and TISE_ERR_INVALID if one is missing / not a regular file);
 *   tise_read_files  file i -> dst + offsets[i] (sizes[i] bytes, as returned above; dst is typically pinned memory);
 *                    on failure *failed_index names a file and the call returns TISE_ERR_INVALID;
 *   tise_probe_many  tise_png_probe, else (use_jpeg) tise_jpeg_probe on every staged file:
 *                    kind[i] = 1 PNG / 2 JPEG the device decoders handle / 0 anything else, hw[i] = {height, width}. */
int tise_file_sizes(const char* const* paths, int n, int64_t* sizes, int n_threads);
int tise_read_files(const char* const* paths, const int64_t* offsets, const int64_t* sizes, int n, uint8_t* dst, int n_threads,
                    int* failed_index);
int tise_probe_many(const uint8_t* h_base, const int64_t* offsets, const int64_t* sizes, int n, int use_jpeg,
                    tise_png_info* png_infos, tise_jpeg_info* jpeg_infos, int8_t* kind, int32_t* hw, int n_threads);
size_t tise_jpeg_decode_workspace_bytes(const tise_jpeg_info* infos, int n);
int tise_jpeg_decode(tise_ctx* ctx, const uint8_t* h_files, const uint8_t* d_files, const int64_t* h_file_off,
                     const int64_t* h_file_bytes, int n, uint8_t* d_out, const int64_t* h_out_off,
                     int32_t* d_status, void* d_ws, size_t ws_bytes, tise_stream stream);

/* ---- A2: Pillow-exact bilinear resize to 299x299 ---------------------------------------- */
/* transforms.Resize((299,299)) at fid_score.py:210 / object_centric_inception_score.py:89 and
 * scipy.misc.imresize at inception_score_star_coco.py:118, inception_score_star_bird.py:67:
 * Pillow's two-pass fixed-point triangle filter, bit-exact.
 * d_imgs: packed uint8 HWC RGB images; image i starts at byte h_offsets[i] and has
 * h_hw[2i] rows x h_hw[2i+1] columns.  If uniform != 0, h_offsets may be NULL and h_hw holds
 * one (h, w) pair for all images.  d_out: n x 299 x 299 x 3 uint8. */
size_t tise_resize_workspace_bytes(int n, int max_h, int max_w);
int tise_resize_u8(tise_ctx* ctx, const uint8_t* d_imgs, const int64_t* h_offsets, const int32_t* h_hw,
                   int uniform, int n, uint8_t* d_out, void* d_ws, size_t ws_bytes, tise_stream stream);

/* Crop hand-off: object_fidelity/crop_object.py:38-46 writes every detected object as
 * img.crop(box) to a PNG which O-FID / O-IS later re-open and resize to 299x299
 * (fid_score.py:210, object_centric_inception_score.py:89).  This entry point does crop + resize in one
 * pass straight from the source images in HBM: crop i = pixels [y0, y1) x [x0, x1) of the image at
 * d_imgs + h_offsets[i] (h_hw[2i] rows x h_hw[2i+1] columns; several crops may name the same image),
 * h_box = n x (x0, y0, x1, y1) already rounded as PIL's Image.crop rounds (int(round(v))).
 * Output bytes are identical to Image.crop(box).resize((299, 299), BILINEAR).  Boxes must be non-empty
 * and inside their image (Detectron2 clips its boxes); otherwise TISE_ERR_INVALID.
 * Workspace: tise_resize_workspace_bytes(n, max box height, max box width). */
int tise_crop_resize_u8(tise_ctx* ctx, const uint8_t* d_imgs, const int64_t* h_offsets, const int32_t* h_hw,
                        const int32_t* h_box, int n, uint8_t* d_out, void* d_ws, size_t ws_bytes,
                        tise_stream stream);

/* Float bilinear resampling of n x c x h x w (contiguous NCHW) to n x c x 299 x 299 for callers that hand the
 * reference-facing modules float tensors of another size: F.upsample(..., align_corners=True) at
 * image_realism/FID/inception.py:117-118 (align_corners = 1) and nn.Upsample(size=(299, 299), mode="bilinear") at
 * object_centric_inception_score.py:49 (align_corners = 0).  Same source-index arithmetic as ATen (float). */
int tise_upsample_bilinear_f32(tise_ctx* ctx, const float* d_in, int n, int c, int h, int w, int align_corners,
                               float* d_out, tise_stream stream);

/* Per-channel affine of a float NCHW tensor (c <= 4, hw = H * W), y = x * scale[c] + bias[c] as a rounded multiply and a
 * rounded add: the normalize_input pre-amble of image_realism/FID/inception.py:119-124 for callers that hand
 * InceptionV3.forward float tensors (the uint8 path carries it in the lookup table of tise_set_input_lut).
 * h_scale / h_bias: host arrays of c floats.  d_out may alias d_in. */
int tise_channel_affine_f32(tise_ctx* ctx, const float* d_in, int n, int c, int hw, const float* h_scale, const float* h_bias,
                            float* d_out, tise_stream stream);

/* ---- A3/A4: uint8 level -> network input ------------------------------------------------ */
/* lut[c*256 + v] = float fed to Conv2d_1a_3x3 for channel c and uint8 level v; covers ToTensor
 * (fid_score.py:211) + the affine at image_realism/FID/inception.py:120-124, or
 * Normalize(.5,.5) (object_centric_inception_score.py:91), or x/127.5-1
 * (inception_score_star_bird.py:70), or the coco graph's (x-128)/128. */
int tise_set_input_lut(tise_ctx* ctx, const float* h_lut768);

/* ---- A5/A6: Inception-v3 trunk ----------------------------------------------------------- */
/* InceptionV3.forward (image_realism/FID/inception.py:100-134) for n images.
 * Exactly one of d_u8 (n x 299 x 299 x 3 uint8, goes through the LUT) and d_f32
 * (n x 3 x 299 x 299 float, already normalised, as the reference's model(batch) receives after
 * its pre-amble) must be non-NULL.  d_feat: n x dims(tap) float32, the (spatially averaged)
 * output of the requested block: TISE_TAP_2048 is pool3. */
size_t tise_inception_workspace_bytes(const tise_ctx* ctx, int n);
int tise_inception_features(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, int tap,
                            float* d_feat, void* d_ws, size_t ws_bytes, tise_stream stream);

/* fc head: d_logits[n x num_classes] = d_feat[n x 2048] . fc.weight^T (+ fc.bias if use_bias).
 * object_centric_inception_score.py:52 (with bias), inception_score_star_coco.py:104-105
 * (no bias), inception_score_star_bird.py:181-186 (with bias). */
int tise_logits(tise_ctx* ctx, const float* d_feat, int n, int use_bias, float* d_logits, tise_stream stream);

/* ---- A7: activation statistics ------------------------------------------------------------ */
/* One-pass replacement of np.mean / np.cov at fid_score.py:194-195: adds this batch's
 * sum_i x_i into d_sum[d] and sum_i x_i x_i^T into the LOWER triangle of d_sxx[d x d]
 * (fp64 accumulation of exact fp32 products; buffers must be zeroed by the caller first). */
int tise_moments_update(tise_ctx* ctx, const float* d_feat, int n, int d, double* d_sum, double* d_sxx,
                        tise_stream stream);
/* mu = sum / count, sigma = (sxx - count mu mu^T) / (count - 1), full symmetric matrix. */
int tise_moments_finalize(tise_ctx* ctx, const double* d_sum, const double* d_sxx, int64_t count, int d,
                          double* d_mu, double* d_sigma, tise_stream stream);

/* ---- A11-A13: inception score partial sums ------------------------------------------------- */
/* For each image i with split s = d_split_id[i] >= 0: p = softmax(logits[i, class_offset:C] / T)
 * (fp32, like the reference), then d_splogp[s] += sum_c p_c log p_c, d_marg[s*Cs + c] += p_c,
 * d_cnt[s] += 1, d_zeros[0] += #(p_c == 0)   (fp64 / int64 accumulators, Cs = C - class_offset).
 * Replaces object_centric_inception_score.py:54-57,72-79,
 * inception_score_star_coco.py:53-58, inception_score_star_bird.py:98-105. */
int tise_is_update(tise_ctx* ctx, const float* d_logits, int n, int C, int class_offset, float temperature,
                   const int32_t* d_split_id, int n_splits, double* d_splogp, double* d_marg,
                   long long* d_cnt, long long* d_zeros, tise_stream stream);

/* ---- A14: temperature calibration ---------------------------------------------------------- */
/* One closure evaluation of classifier_calibration/temperature_scaling.py:56-63 plus the ECE of :97-119 on
 * n x C validation logits and int64 labels:  accumulates into d_out (double, zeroed by the caller,
 * 4 + 3 n_bins + 1 entries)
 *   [0] sum_i nll_i of CrossEntropyLoss(logits / T, labels)      [1] sum_i d nll_i / dT  (closed form)
 *   [2] rows                                                       [3] correct top-1 predictions
 *   [4 ..) per confidence bin (h_bin_bounds[b] < conf <= h_bin_bounds[b+1], n_bins + 1 bounds as
 *   torch.linspace(0, 1, n_bins + 1) gives them): count, sum of confidences, sum of accuracies
 *   [4 + 3 n_bins] an int32 flag, non-zero if a label was outside [0, C).
 * The LBFGS iteration itself is a scalar recurrence and stays with the caller. */
int tise_calibration_eval(tise_ctx* ctx, const float* d_logits, const int64_t* d_labels, int n, int C,
                          float temperature, const float* h_bin_bounds, int n_bins, double* d_out,
                          tise_stream stream);

/* ---- A9: Frechet distance ------------------------------------------------------------------- */
/* calculate_frechet_distance (fid_score.py:121-171):
 *   |mu1-mu2|^2 + tr(s1) + tr(s2) - 2 tr sqrt(s1 s2),
 * with tr sqrt(s1 s2) evaluated in fp64 on the device as the sum of square roots of the
 * eigenvalues of the symmetric matrix L^T s2 L (s1 + delta I = L L^T), which equal those of s1 s2.
 * Writes the scalar to *h_fid (host) and synchronises the stream.  flags bit0: a non-finite
 * value appeared and the eps-regularised retry of fid_score.py:156-160 was taken. */
size_t tise_frechet_workspace_bytes(int d);
int tise_frechet(tise_ctx* ctx, const double* d_mu1, const double* d_sigma1, const double* d_mu2,
                 const double* d_sigma2, int d, double eps, double* h_fid, int* h_flags, void* d_ws,
                 size_t ws_bytes, tise_stream stream);

/* ---- diagnostics ------------------------------------------------------------------------------ */
/* One Conv2d(bias=False)+scale/shift+ReLU layer through the selected engine, used by the tests to
 * check the UMMA kernels shape by shape.  x: n x h x w x cin fp16 NHWC (device); w: cout x cin x kh x kw
 * fp32 (host); scale/shift: cout fp32 (host); y: n x ho x wo x cout fp16 NHWC (device).
 * Synchronises the device. */
int tise_debug_conv(tise_ctx* ctx, int impl, int passes, const void* d_x, int n, int h, int w, int cin,
                    const float* h_w, const float* h_scale, const float* h_shift, int cout, int kh, int kw,
                    int stride, int pad_h, int pad_w, int relu, void* d_y);

/* Per-kernel-kind device timing of tise_inception_features, taken with CUDA events recorded on the
 * launch stream around every kernel (kinds: 0 stem conv, 1 tcgen05/SIMT conv, 2 max pool, 3 avg pool,
 * 4 global avg pool).  Enabling clears earlier records; tise_profile_read waits for the recorded events
 * and fills three arrays of 8: total milliseconds, launches, and algorithmic FLOPs (2 M N K of the
 * convolutions, padding channels excluded).  bench.py derives roofline.achieved from these. */
int tise_profile_enable(tise_ctx* ctx, int on);
int tise_profile_read(tise_ctx* ctx, double* ms_by_kind, long long* launches_by_kind, double* flops_by_kind);

/* Number of kernels this library has launched on this ctx since creation (for bench.py's gpu_launches). */
long long tise_launch_count(const tise_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TISE_B200_H */
