// extern "C" doorway into the UNMODIFIED reference C++ golden model.
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
// /root/reference/hw/sim/cpp/src/golden_models.cpp (from where it lies; no reference
// source is copied into this repo) into oracle/_ref/libref_golden.so.
//
// Every function below only forwards to a reference function; the functions that the
// reference defines in golden_models.cpp but does not declare in golden_models.hpp
// are forward-declared here with the signatures of their definitions.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "bsr_packer.hpp"
#include "golden_models.hpp"

namespace resnet_accel {
namespace golden {
// defined at golden_models.cpp:465, :534, :601, :684, :801, :883 (not all declared in the header)
void add_residual_int8(const std::int8_t*, const std::int8_t*, std::int8_t*, std::size_t, float, float, float);
void maxpool2d_int8(const std::int8_t*, std::int8_t*, std::size_t, std::size_t, std::size_t, std::size_t, std::size_t);
void avgpool_global_int8(const std::int8_t*, std::int8_t*, std::size_t, std::size_t, std::size_t);
void conv2d_int8_simple(const std::int8_t*, const std::int8_t*, const std::int32_t*, std::int32_t*, std::size_t,
                        std::size_t, std::size_t, std::size_t, std::size_t, std::size_t, std::size_t);
void im2col_int8(const std::int8_t*, std::int8_t*, std::size_t, std::size_t, std::size_t, std::size_t, std::size_t,
                 std::size_t, std::size_t, std::size_t);
void conv2d_int8_im2col(const std::int8_t*, const std::int8_t*, const std::int32_t*, std::int32_t*, std::size_t,
                        std::size_t, std::size_t, std::size_t, std::size_t, std::size_t, std::size_t);
}  // namespace golden
}  // namespace resnet_accel

namespace g = resnet_accel::golden;
#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_matmul_int8(const int8_t* A, const int8_t* B, int32_t* C, size_t M, size_t K, size_t N) {
  g::matmul_int8(A, B, C, M, K, N);
}

// Convention A: dense B[K,N] is packed by the reference's own pack_to_bsr, then multiplied by
// the reference's bsr_matmul_int8.  Returns the number of stored blocks (or -1 if validate_bsr fails).
REF_API long ref_pack_and_bsr_matmul_int8(const int8_t* A, const int8_t* B_dense, int32_t* C, size_t M, size_t K,
                                          size_t N) {
  resnet_accel::BSRMatrix bsr = resnet_accel::pack_to_bsr(B_dense, K, N);
  if (!resnet_accel::validate_bsr(bsr).valid) return -1;
  g::bsr_matmul_int8(A, bsr, C, M, K, N);
  return static_cast<long>(bsr.nnz_blocks);
}

// pack_to_bsr structure read-out: caller passes buffers sized for the worst case.
REF_API long ref_pack_to_bsr(const int8_t* dense, size_t rows, size_t cols, int64_t* row_ptr, int64_t* col_idx,
                             int8_t* data) {
  resnet_accel::BSRMatrix bsr = resnet_accel::pack_to_bsr(dense, rows, cols);
  for (size_t i = 0; i < bsr.row_ptr.size(); ++i) row_ptr[i] = static_cast<int64_t>(bsr.row_ptr[i]);
  for (size_t i = 0; i < bsr.col_idx.size(); ++i) col_idx[i] = static_cast<int64_t>(bsr.col_idx[i]);
  std::memcpy(data, bsr.data.data(), bsr.data.size());
  return static_cast<long>(bsr.nnz_blocks);
}

REF_API void ref_relu_int8(int8_t* d, size_t n) { g::relu_int8(d, n); }
REF_API void ref_relu_int32(int32_t* d, size_t n) { g::relu_int32(d, n); }
REF_API void ref_relu6_int8(int8_t* d, size_t n, float scale) { g::relu6_int8(d, n, scale); }

REF_API void ref_requantize_int32_to_int8(const int32_t* in, int8_t* out, size_t n, float in_scale, float out_scale) {
  g::requantize_int32_to_int8(in, out, n, in_scale, out_scale);
}

REF_API void ref_add_residual_int8(const int8_t* a, const int8_t* b, int8_t* out, size_t n, float sa, float sb,
                                   float so) {
  g::add_residual_int8(a, b, out, n, sa, sb, so);
}

REF_API void ref_maxpool2d_int8(const int8_t* in, int8_t* out, size_t H, size_t W, size_t C, size_t pool,
                                size_t stride) {
  g::maxpool2d_int8(in, out, H, W, C, pool, stride);
}

REF_API void ref_avgpool_global_int8(const int8_t* in, int8_t* out, size_t H, size_t W, size_t C) {
  g::avgpool_global_int8(in, out, H, W, C);
}

REF_API void ref_im2col_int8(const int8_t* in, int8_t* col, size_t C, size_t H, size_t W, size_t K, size_t stride,
                             size_t pad, size_t Ho, size_t Wo) {
  g::im2col_int8(in, col, C, H, W, K, stride, pad, Ho, Wo);
}

REF_API void ref_conv2d_int8_simple(const int8_t* in, const int8_t* w, const int32_t* bias, int32_t* out, size_t Cin,
                                    size_t H, size_t W, size_t Cout, size_t K, size_t stride, size_t pad) {
  g::conv2d_int8_simple(in, w, bias, out, Cin, H, W, Cout, K, stride, pad);
}

REF_API void ref_conv2d_int8_im2col(const int8_t* in, const int8_t* w, const int32_t* bias, int32_t* out, size_t Cin,
                                    size_t H, size_t W, size_t Cout, size_t K, size_t stride, size_t pad) {
  g::conv2d_int8_im2col(in, w, bias, out, Cin, H, W, Cout, K, stride, pad);
}

// The reference's conv-layer recipe for ONE CHW image, composed only of reference calls:
// conv2d_int8_im2col (dense OIHW weights, INT32 bias) -> relu_int32 -> requantize_int32_to_int8
// once per output-channel plane with in_scale = s_act*s_w[c] (float32 product formed by the caller)
// -> add_residual_int8.  `scratch` holds Cout*Ho*Wo int32.
REF_API void ref_conv_layer_image(const int8_t* in, const int8_t* w, const int32_t* bias, size_t Cin, size_t H,
                                  size_t W, size_t Cout, size_t K, size_t stride, size_t pad, int relu,
                                  const float* in_scale_per_channel, float out_scale, const int8_t* residual,
                                  float s_main, float s_res, float s_out, int32_t* scratch, int8_t* out) {
  const size_t Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1, P = Ho * Wo;
  g::conv2d_int8_im2col(in, w, bias, scratch, Cin, H, W, Cout, K, stride, pad);
  if (relu) g::relu_int32(scratch, Cout * P);
  for (size_t c = 0; c < Cout; ++c)
    g::requantize_int32_to_int8(scratch + c * P, out + c * P, P, in_scale_per_channel[c], out_scale);
  if (residual) g::add_residual_int8(out, residual, out, Cout * P, s_main, s_res, s_out);
}

// serialize_for_hardware / deserialize_from_hardware (bsr_packer.hpp:489-575) on the reference's own pack_to_bsr of a dense
// int8 matrix.  Returns the blob size (or -1 if `cap` is too small); round-trips it through the reference's reader and
// returns -2 if verify_serialization fails.
REF_API long ref_serialize_for_hardware(const int8_t* dense, size_t rows, size_t cols, uint8_t* out, size_t cap) {
  resnet_accel::BSRMatrix bsr = resnet_accel::pack_to_bsr(dense, rows, cols);
  std::vector<std::uint8_t> blob = resnet_accel::serialize_for_hardware(bsr);
  if (!resnet_accel::verify_serialization(bsr)) return -2;
  if (blob.size() > cap) return -1;
  std::memcpy(out, blob.data(), blob.size());
  return static_cast<long>(blob.size());
}
// The reference's reader on a caller-supplied blob: fills the structure arrays, returns nnz (or -1 when it throws).
REF_API long ref_deserialize_from_hardware(const uint8_t* blob, size_t size, int64_t* hdr3, int64_t* row_ptr, int64_t* col_idx,
                                           int8_t* data) {
  try {
    resnet_accel::BSRMatrix bsr = resnet_accel::deserialize_from_hardware(blob, size);
    hdr3[0] = static_cast<int64_t>(bsr.nnz_blocks); hdr3[1] = static_cast<int64_t>(bsr.num_block_rows);
    hdr3[2] = static_cast<int64_t>(bsr.num_block_cols);
    for (size_t i = 0; i < bsr.row_ptr.size(); ++i) row_ptr[i] = static_cast<int64_t>(bsr.row_ptr[i]);
    for (size_t i = 0; i < bsr.col_idx.size(); ++i) col_idx[i] = static_cast<int64_t>(bsr.col_idx[i]);
    std::memcpy(data, bsr.data.data(), bsr.data.size());
    return static_cast<long>(bsr.nnz_blocks);
  } catch (const std::exception&) {
    return -1;
  }
}
