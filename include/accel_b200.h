/*
 * accel_b200.h - C ABI of the B200-native ACCEL-v1 hot path (libaccel_b200.so).
 *
 * Drop-in boundary for the reference's BSR-INT8 path.  The reference itself has no FFI
 * (SURVEY.md 8b): its boundary is a set of Python call signatures plus the C++ host class.
 * Every entry point below names the reference interface it stands in for (file:line under
 * the reference root).  All `*_dev` / unqualified data pointers are DEVICE pointers owned by
 * the caller (PyTorch allocates; this library never allocates device memory); `*_host`
 * pointers are host pointers.  Functions return 0 or a negative accel_status that mirrors
 * AcceleratorError::Code (hw/sim/cpp/include/accelerator_driver.hpp:337-345).
 * Calls are asynchronous on `stream` unless stated otherwise.  No CPU fallback exists.
 */
#ifndef ACCEL_B200_H
#define ACCEL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ACCEL_API __attribute__((visibility("default")))
#else
#define ACCEL_API
#endif

typedef void* accel_stream_t; /* cudaStream_t */

typedef enum accel_status {
  ACCEL_OK = 0,
  ACCEL_INIT_FAILED = -1,     /* AcceleratorError::INIT_FAILED    */
  ACCEL_TIMEOUT = -2,         /* AcceleratorError::TIMEOUT        */
  ACCEL_DMA_ERROR = -3,       /* AcceleratorError::DMA_ERROR      (CUDA runtime error) */
  ACCEL_ILLEGAL_COMMAND = -4, /* AcceleratorError::ILLEGAL_COMMAND */
  ACCEL_INVALID_CONFIG = -5,  /* AcceleratorError::INVALID_CONFIG (bad sizes / BSR structure) */
  ACCEL_MEMORY_ERROR = -6,    /* AcceleratorError::MEMORY_ERROR   (workspace too small / misaligned) */
  ACCEL_NOT_READY = -7        /* AcceleratorError::NOT_READY      (plan not uploaded) */
} accel_status;

/* Epilogue flags (golden_models.cpp:298-303, :378-411, :465-490). */
enum {
  ACCEL_RELU = 1 << 0,      /* relu_int32 on the accumulator before scaling            */
  ACCEL_OUT_I8 = 1 << 1,    /* per-channel requant -> int8 (needs chan_scale)          */
  ACCEL_OUT_I32 = 1 << 2,   /* raw INT32 accumulator (+bias, +relu)                    */
  ACCEL_OUT_F32 = 1 << 3,   /* float32(acc) * chan_scale[c]  (de-quantised logits)     */
  ACCEL_RELU_OUT = 1 << 4   /* relu_int8 on the final int8 value (after the residual add, golden_models.cpp:278) */
};

/* Where activation row m, output channel c lands (pix = m % rows_per_image):
 *   row_len == 0:  out[(m / rows_per_image) * image_stride + c * chan_stride + pix * row_stride]
 *   row_len  > 0:  out[(m / rows_per_image) * image_stride + c * chan_stride
 *                      + (pix / row_len) * row_pitch + (pix % row_len) * row_stride]
 * Row-major [M, ld]:  rows_per_image = M, image_stride = 0, chan_stride = 1, row_stride = ld.
 * NCHW conv output:   rows_per_image = Ho*Wo, image_stride = Cout*Ho*Wo, chan_stride = Ho*Wo, row_stride = 1.
 * NCHW with padded rows (pitch Wp >= Wo, e.g. a multiple of 16 so that the next layer can stream rows with
 * 16-byte copies): row_len = Wo, row_pitch = Wp, chan_stride = Ho*Wp, image_stride = Cout*Ho*Wp. */
typedef struct accel_out_layout {
  int64_t rows_per_image;
  int64_t image_stride;
  int64_t chan_stride;
  int64_t row_stride;
  int64_t row_len;
  int64_t row_pitch;
} accel_out_layout;

typedef struct accel_epilogue {
  int32_t flags;             /* ACCEL_RELU | one ACCEL_OUT_*                                   */
  int32_t n_channels;        /* channels actually written (<= n_block_rows*14)                  */
  const float* chan_scale;   /* [n_channels] float32 requant factor sf[c]; may be NULL for I32  */
  const int32_t* bias;       /* [n_channels] INT32 bias added to the accumulator; may be NULL   */
  const int8_t* residual;    /* same layout as out (int8); NULL = no residual add               */
  float res_scale_main;      /* add_residual_int8(main, res, s_main, s_res, s_out)              */
  float res_scale_res;
  float res_scale_out;
  unsigned long long* sat_count; /* device counter += #values clipped by the int8 requant; may be NULL */
  int32_t* chan_absmax;      /* device [n_channels], atomicMax of |acc| per channel; may be NULL */
  int32_t acc_bound;         /* optional promise by the caller: |accumulator + bias| <= acc_bound for EVERY possible int8 input
                                (128 * the largest row L1 norm of the INT8 weights + the largest |bias|).  Below 2^22 the
                                weight-stationary kernels then run their conversion-free epilogue (same results bit for bit);
                                0 = unknown, the general epilogue. */
} accel_epilogue;

/* Convolution geometry: input int8 NCHW, weights viewed [Cout, Cin*k*k] with K index
 * (c_in, kh, kw) exactly as im2col_int8 orders it (golden_models.cpp:801-842). */
typedef struct accel_conv_geom {
  int32_t batch, c_in, h, w;
  int32_t ksize, stride, pad;
  int32_t in_row_pitch; /* bytes between input rows (>= w); 0 = w.  Planes / images are dense in that pitch. */
} accel_conv_geom;

typedef struct accel_plan accel_plan; /* opaque: host-side schedule of one BSR weight matrix */

/* --- library ---------------------------------------------------------------------------- */
ACCEL_API const char* accel_last_error_string(void);
ACCEL_API const char* accel_version(void);
/* 0 if a sm_100 device is current and the tcgen05 path can run, else ACCEL_INIT_FAILED. */
ACCEL_API int accel_device_check(void);
/* Developer aid: while non-NULL, every tensor-core launch writes 8 clock64 stamps per CTA into `dev_buffer`
 * (kernel entry, prologue done, first activation stage acquired / published, main loop done by the producers,
 * accumulators complete, epilogue done, CTA exit).  The buffer must hold 8 int64 per CTA of the largest launch. */
ACCEL_API void accel_debug_set_timeline(long long* dev_buffer);
/* Developer aid: event counters of this process.  which = 0: launches of the weight-stationary convolution kernels,
 * 1: launches of the dense-equivalent GEMM kernel (gemm_ws_kernel), 2: conv launches that ran the conversion-free epilogue
 * (accel_epilogue.acc_bound). */
ACCEL_API long long accel_debug_counter(int which);

/* --- weight plan: replaces AccelDriver.load_sparse_weights (sw/host/accel.py:177-236) and
 *     AcceleratorDriver::set_layer_weights / load_weights_bsr (accelerator_driver.cpp:643-760).
 * Validates the structure with the rules of validate_bsr (bsr_packer.hpp:364-436), builds the
 * MMA schedule on the host and reports the device workspace it needs.  Convention B:
 * block-rows = output channels, col_idx = K tile.  block must be 14 (the tensor-core path). */
ACCEL_API int accel_plan_create(const int32_t* row_ptr_host, const int32_t* col_idx_host, int32_t n_block_rows,
                      int32_t n_block_cols, int32_t block, int32_t group_rows_hint /* 0 = default (32) */,
                      accel_plan** plan_out, size_t* workspace_bytes);
/* Repack the 14x14 blocks (`blocks_dev`, nnz*196 int8, reference layout) into 16-aligned MMA
 * tiles inside `workspace_dev` (>= workspace_bytes, 256-byte aligned) and upload the schedule. */
ACCEL_API int accel_plan_upload(accel_plan* plan, const int8_t* blocks_dev, void* workspace_dev, size_t workspace_bytes,
                      accel_stream_t stream);
ACCEL_API void accel_plan_destroy(accel_plan* plan);
ACCEL_API int64_t accel_plan_num_blocks(const accel_plan* plan);
ACCEL_API int64_t accel_plan_num_mma(const accel_plan* plan);   /* tcgen05.mma instructions per 128-row tile */
ACCEL_API int64_t accel_plan_num_tiles(const accel_plan* plan); /* 16x32 weight tiles (512 B each) */
/* Schedule read-out for tests / tooling.  export_ops: one record of 8 int32 per weight tile, in blob order,
 * {group, first_block_row, local_row, k_chunk, window_tile, block_lo, block_hi, batch}; returns the tile count.
 * export_mma: one record of 8 int32 per tcgen05.mma {group, batch, accumulator_column, N, activation_column,
 * first_tile (blob order), k_chunk, first_block_row}; returns the MMA count. */
ACCEL_API int64_t accel_plan_export_ops(const accel_plan* plan, int32_t* records, int64_t cap);
ACCEL_API int64_t accel_plan_export_mma(const accel_plan* plan, int32_t* records, int64_t cap);

/* --- K1+K3: block-sparse GEMM, tcgen05 kind::i8, fused epilogue.
 * Replaces gemm_bsr_int8_golden (sw/golden/golden_fc1_test.py:49-108), AccelDriver.run_inference
 * (sw/host/accel.py:279-342) and AcceleratorDriver::run_layer (accelerator_driver.cpp:435-496).
 * act: int8 [M, K] row-major with leading dimension lda (K may be the unpadded width). */
ACCEL_API int accel_bsr_gemm_i8(const accel_plan* plan, const int8_t* act, int64_t M, int64_t K, int64_t lda,
                      const accel_epilogue* epi, void* out, const accel_out_layout* layout, accel_stream_t stream);

/* --- K2+K1+K3: implicit-im2col BSR convolution (conv2d_int8_im2col, golden_models.cpp:883-933,
 * with the BSR weights of export_bsr_14x14.py).  Output NCHW through `layout`. */
ACCEL_API int accel_conv_bsr_i8(const accel_plan* plan, const int8_t* input_nchw, const accel_conv_geom* geom,
                      const accel_epilogue* epi, void* out, const accel_out_layout* layout, accel_stream_t stream);

/* --- a 3x3 / stride 2 / pad 1 convolution and the 1x1 / stride 2 / pad 0 convolution of the SAME input in one call: the
 * first convolution of a ResNet stage and its downsample branch (two run_layer calls in the reference,
 * resnet_inference.cpp:61-127; the downsample convolutions are named in export_resnet18_bsr.py:72,79,86).  Both outputs
 * share `layout`.  When both plans carry a weight-stationary layout (below) the 1x1 convolution rides along as one more
 * tap of the staged activation tiles; otherwise this is the two accel_conv_bsr_i8 calls in sequence. */
ACCEL_API int accel_conv_bsr_i8_dual(const accel_plan* plan, const accel_plan* plan_ds, const int8_t* input_nchw,
                           const accel_conv_geom* geom, const accel_epilogue* epi, void* out, const accel_epilogue* epi_ds,
                           void* out_ds, const accel_out_layout* layout, accel_stream_t stream);

/* --- the ResNet stem in one call: 7x7 / stride 2 / pad 3 convolution (conv2d_int8_im2col, golden_models.cpp:883-933) with
 * ReLU + per-channel requant (:298-303, :378-411) followed by the 3x3 / stride 2 / pad 1 max-pool (maxpool2d_int8 :534-571,
 * padding = -128), layer table resnet_inference.cpp:61-127.  Writes only the pooled tensor int8 [batch, c_out, H/4, W/4]
 * with rows of `out_pitch` bytes (0 = dense).  The pool is taken on the INT32 accumulators, which is bit-identical because
 * requant is monotone for the positive per-channel factors the exporters produce (the caller guarantees chan_scale > 0);
 * epi->sat_count still counts every clipped convolution output.  Needs a plan prepared with
 * accel_plan_conv_ws_prepare(ksize = 7); returns ACCEL_ILLEGAL_COMMAND when the geometry has no fused kernel (then run
 * accel_conv_bsr_i8 and accel_maxpool_i8). */
ACCEL_API int accel_conv_pool_bsr_i8(const accel_plan* plan, const int8_t* input_nchw, const accel_conv_geom* geom,
                           const accel_epilogue* epi, int32_t pool, int32_t pool_stride, int32_t pool_pad, int8_t* out,
                           int32_t out_pitch, accel_stream_t stream);

/* --- weight-stationary re-layout for 3x3 stride-1 pad-1 convolutions (conv2d_int8_im2col, golden_models.cpp:883-933,
 * K order (c_in, kh, kw) of im2col_int8 :801-842).  Optional: when a plan has been prepared for (c_in, c_out) and the
 * tensors of an accel_conv_bsr_i8 call allow it (16-byte aligned rows, int8 output, width <= 62; ksize 3 with stride 1
 * or 2, ksize 1 as the second plan of accel_conv_bsr_i8_dual, ksize 7 with c_in <= 4 and c_out <= 64 for
 * accel_conv_pool_bsr_i8), that call runs the
 * kernel of csrc/conv_ws.cuh: the stored blocks are scattered once into 128x32 K-major weight tiles per
 * (channel group, 32-channel chunk, tap) and the activation tile is fed to the tensor core straight from TMA.
 * conv_ws_bytes reports the workspace for that layout (0 = no such path for this geometry); conv_ws_prepare fills it
 * (`workspace_dev` 1024-byte aligned, owned by the caller for the life of the plan; synchronises `stream`). */
ACCEL_API int accel_plan_conv_ws_bytes(const accel_plan* plan, int32_t c_in, int32_t c_out, int32_t ksize, size_t* bytes);
ACCEL_API int accel_plan_conv_ws_prepare(accel_plan* plan, const int8_t* blocks_dev, int32_t c_in, int32_t c_out, int32_t ksize,
                               void* workspace_dev, size_t workspace_bytes, accel_stream_t stream);
/* Forget the prepared layout (call before freeing or reusing its workspace): later convolutions take the gather kernels. */
ACCEL_API void accel_plan_conv_ws_release(accel_plan* plan);

/* --- dense-equivalent layout for GEMMs (csrc/gemm_ws.cuh).  Optional: when a plan has been prepared and the tensors of an
 * accel_bsr_gemm_i8 call allow it (16-byte aligned activation rows, a strided output layout), that call runs
 * gemm_ws_kernel: the stored blocks are scattered once into a plain row-major int8 matrix [channels, K] (neither padded
 * 14 -> 16), both operands reach the tensor core by TMA, CTA pairs share 256 x 256 tiles (tcgen05.mma.cta_group::2), and
 * (channel tile, 128-k chunk) regions without any non-zero weight are skipped.  Same results as the gather kernel, bit for
 * bit (gemm_bsr_int8_golden, sw/golden/golden_fc1_test.py:49-108).  gemm_ws_bytes reports the workspace (0 = no such
 * layout for this plan); gemm_ws_prepare fills it (`workspace_dev` 256-byte aligned, owned by the caller until
 * gemm_ws_release or plan destruction; synchronises `stream`).  live_chunks: number of (tile, chunk) regions the kernel
 * visits per 256-row activation tile, for cta_group 1 or 2 (tooling). */
ACCEL_API int accel_plan_gemm_ws_bytes(const accel_plan* plan, size_t* bytes);
ACCEL_API int accel_plan_gemm_ws_prepare(accel_plan* plan, const int8_t* blocks_dev, void* workspace_dev, size_t workspace_bytes,
                               accel_stream_t stream);
ACCEL_API void accel_plan_gemm_ws_release(accel_plan* plan);
ACCEL_API int64_t accel_plan_gemm_ws_live_chunks(const accel_plan* plan, int32_t cta_group);

/* --- reference-shaped CUDA-core kernels: any block size (4/8/14/16 fixtures), Convention B or A.
 * Same arithmetic as above, used for the generic-block path and as an on-device cross-check.
 * orient 0 = Convention B (golden_fc1_test.py), 1 = Convention A (golden_models.cpp:187-255). */
ACCEL_API int accel_bsr_gemm_generic(const int8_t* act, int64_t M, int64_t K, int64_t lda, const int32_t* row_ptr,
                           const int32_t* col_idx, const int8_t* blocks, int32_t n_block_rows, int32_t block_h,
                           int32_t block_w, int32_t orient, int64_t n_out, int32_t* out, int64_t ldo,
                           accel_stream_t stream);

/* --- K4: GPU packer / pruner (export_bsr_14x14.py:406-484, export_bsr.py:76-153,
 * blocksparse_train.py:93-138).  Two phases so that the caller owns every allocation. */
/* per-block statistics of a dense [rows, cols] matrix: int8 -> L1 sums (int32), float -> L2 norms (float32) */
ACCEL_API int accel_block_l1_i8(const int8_t* w, int64_t rows, int64_t cols, int64_t ld, int32_t block, int32_t* l1_out,
                      accel_stream_t stream);
ACCEL_API int accel_block_l2_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, int32_t bh, int32_t bw, float* l2_out,
                       accel_stream_t stream);
/* keep[i] (uint8, [nbr*nbc]) -> row_ptr [nbr+1] (int32) and per-block output slot (int32, -1 = dropped) */
ACCEL_API int accel_bsr_scan(const uint8_t* keep, int32_t nbr, int32_t nbc, int32_t* row_ptr, int32_t* slot,
                   accel_stream_t stream);
/* gather kept blocks: col_idx [nnz], blocks [nnz, block, block] (zero padded at the edges) */
ACCEL_API int accel_bsr_gather_i8(const int8_t* w, int64_t rows, int64_t cols, int64_t ld, int32_t block, const int32_t* slot,
                        int32_t nbr, int32_t nbc, int32_t* col_idx, int8_t* blocks, accel_stream_t stream);
ACCEL_API int accel_bsr_gather_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, int32_t bh, int32_t bw,
                                   const int32_t* slot, int32_t nbr, int32_t nbc, int32_t* col_idx, float* blocks,
                                   accel_stream_t stream);
/* per-row symmetric quantisation q = clip(rint(w / scale[row]), -128, 127) (quantize.py:71-98) */
ACCEL_API int accel_quantize_rows_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, const float* scales, int8_t* q,
                            accel_stream_t stream);
/* scales[i] = max(absmax[i] / 127, 1e-12) in float32 (quantize.py:86) */
ACCEL_API int accel_symmetric_scales_f32(const float* absmax, int64_t n, float* scales, accel_stream_t stream);
ACCEL_API int accel_row_absmax_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, float* absmax,
                         accel_stream_t stream);

/* --- K5 + stand-alone epilogue pieces (golden_models.cpp:378-411, :465-490, :534-628) */
ACCEL_API int accel_requant_i32_i8(const int32_t* acc, int8_t* out, int64_t n_outer, int64_t n_chan, int64_t n_inner,
                         const float* chan_scale, const int32_t* bias, int32_t relu, unsigned long long* sat_count,
                         accel_stream_t stream);
ACCEL_API int accel_add_residual_i8(const int8_t* main_, const int8_t* res, int8_t* out, int64_t n, float s_main, float s_res,
                          float s_out, accel_stream_t stream);
/* relu_int8 / relu6_int8 / relu_int32, in place (golden_models.cpp:278-283, :323-330, :298-303).  relu6: upper clamp
 * int8(6.0f / scale), truncated as the reference computes it. */
/* out[pl][y][x] = in[pl][2y][2x] for `planes` int8 image planes of h x w (rows in_pitch / out_pitch bytes apart): the
 * pixels a 1x1 / stride 2 / pad 0 convolution reads (conv2d_int8_im2col, golden_models.cpp:883-933, stride handling;
 * ResNet-50 downsample branches), so that the convolution itself runs as a stride-1 pointwise one.  Output planes are
 * ceil(h/2) x ceil(w/2); bytes of an output row beyond that are left alone. */
ACCEL_API int accel_subsample2_i8(const int8_t* in, int64_t planes, int32_t h, int32_t w, int32_t in_pitch, int8_t* out,
                                  int32_t out_pitch, accel_stream_t stream);
ACCEL_API int accel_relu_i8(int8_t* data, int64_t n, accel_stream_t stream);
ACCEL_API int accel_relu6_i8(int8_t* data, int64_t n, float scale, accel_stream_t stream);
ACCEL_API int accel_relu_i32(int32_t* data, int64_t n, accel_stream_t stream);
/* gemm_bsr_int8 (sw/golden/gemm_bsr_int8.py:16-104): the reference's float32 "golden" with per-row scales, replayed
 * operation by operation in its (block-row, block, local-row) order - SURVEY.md A.2 lists its quirks; it is a compatibility
 * path, not the numerical oracle.  A int8 [M, K] (lda); BSR of B[K, N] with square blocks: indptr / indices / data [nnz, b, b]
 * and blk_row[nnz] = block-row of every stored block.  NumPy's promotion rules decide the arithmetic dtype and are passed in:
 *   - the re-quantisation  clip(rint(block / scale))  (:74-79) runs in float64 when the blocks OR the scales are float64:
 *     div_f64 says so, `data` and `scales_div` [n_scales] are both in that dtype;
 *   - the de-quantising products and the += (:94-102): scale_a_f64 (scale_A is a float64 NumPy scalar; a Python float is
 *     "weak" and multiplies in float32) and scales_f64 (dtype of `scales` [n_scales], the caller's own array).
 * q_scratch: nnz*b*b bytes.  C float32 [M, N] (ldc), overwritten. */
ACCEL_API int accel_gemm_bsr_int8_fp32(const int8_t* A, int64_t M, int64_t lda, const int32_t* indptr, const int32_t* indices,
                             const void* data, const void* scales_div, int32_t div_f64, const int32_t* blk_row, int64_t nnz,
                             int32_t n_block_rows, int32_t block, int64_t K, int64_t N, double scale_a, int32_t scale_a_f64,
                             const void* scales, int32_t scales_f64, int32_t n_scales, int8_t* q_scratch, float* C, int64_t ldc,
                             accel_stream_t stream);
/* in_pitch / out_pitch: bytes between rows of the input / output planes (0 = dense) */
ACCEL_API int accel_maxpool_i8(const int8_t* x, int8_t* out, int64_t n_planes, int32_t h, int32_t w, int32_t pool,
                     int32_t stride, int32_t pad, int32_t in_pitch, int32_t out_pitch, accel_stream_t stream);
ACCEL_API int accel_avgpool_i8(const int8_t* x, int8_t* out, int64_t n_planes, int32_t h, int32_t w, int32_t in_pitch,
                     accel_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ACCEL_B200_H */
