// CUDA-core kernels: generic-block BSR GEMM (both conventions), GPU packer / pruner pieces,
// stand-alone epilogue pieces and pools.  HBM-bound byte/integer work: coalesced, grid-stride,
// sized in multiples of the SM count by the launchers in api.cu.
#pragma once
#include <cstdint>

namespace accel {

// ------------------------------------------------------------------ generic BSR GEMM (any block)
// Convention B (sw/golden/golden_fc1_test.py:78-106): one thread per (m, output column n).
__global__ void bsr_gemm_generic_b_kernel(const int8_t* __restrict__ X, int64_t M, int64_t K, int64_t lda,
                                          const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                          const int8_t* __restrict__ blocks, int32_t nbr, int32_t bh, int32_t bw,
                                          int64_t n_out, int32_t* __restrict__ Y, int64_t ldo) {
  const int64_t total = M * n_out;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = idx / n_out, n = idx - m * n_out;
    const int32_t br = static_cast<int32_t>(n / bh), h = static_cast<int32_t>(n - static_cast<int64_t>(br) * bh);
    uint32_t acc = 0;  // wraps like the hardware accumulator (mac8.sv:392-399)
    if (br < nbr) {
      const int8_t* x = X + m * lda;
      for (int32_t j = row_ptr[br]; j < row_ptr[br + 1]; ++j) {
        const int64_t k0 = static_cast<int64_t>(col_idx[j]) * bw;
        const int8_t* wrow = blocks + (static_cast<int64_t>(j) * bh + h) * bw;
        const int lim = static_cast<int>(min(static_cast<int64_t>(bw), K - k0));
        int32_t s = 0;
        for (int w = 0; w < lim; ++w) s += static_cast<int32_t>(x[k0 + w]) * static_cast<int32_t>(wrow[w]);
        acc += static_cast<uint32_t>(s);
      }
    }
    Y[m * ldo + n] = static_cast<int32_t>(acc);
  }
}

// Convention A (hw/sim/cpp/src/golden_models.cpp:187-255): B[K, N], block-rows over K, col_idx = N tile.
__global__ void bsr_gemm_generic_a_kernel(const int8_t* __restrict__ A, int64_t M, int64_t K, int64_t lda,
                                          const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_idx,
                                          const int8_t* __restrict__ blocks, int32_t nbr, int32_t bh, int32_t bw,
                                          int64_t N, int32_t* __restrict__ C, int64_t ldo) {
  const int64_t total = M * N;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = idx / N, n = idx - m * N;
    const int32_t bc = static_cast<int32_t>(n / bw), jj = static_cast<int32_t>(n - static_cast<int64_t>(bc) * bw);
    uint32_t acc = 0;
    for (int32_t br = 0; br < nbr; ++br) {
      int32_t lo = row_ptr[br], hi = row_ptr[br + 1];
      while (lo < hi) {  // col_idx is strictly increasing inside a block-row (validate_bsr)
        const int32_t mid = (lo + hi) >> 1;
        if (col_idx[mid] < bc) lo = mid + 1; else hi = mid;
      }
      if (lo < row_ptr[br + 1] && col_idx[lo] == bc) {
        const int8_t* blk = blocks + static_cast<int64_t>(lo) * bh * bw;
        int32_t s = 0;
        for (int i = 0; i < bh; ++i) {
          const int64_t k = static_cast<int64_t>(br) * bh + i;
          if (k < K) s += static_cast<int32_t>(A[m * lda + k]) * static_cast<int32_t>(blk[i * bw + jj]);
        }
        acc += static_cast<uint32_t>(s);
      }
    }
    C[m * ldo + n] = static_cast<int32_t>(acc);
  }
}

// ------------------------------------------------------------------ packer: block statistics
// One warp per block; lanes stride over the block's elements (rows of `bw` contiguous bytes).
__global__ void block_l1_i8_kernel(const int8_t* __restrict__ w, int64_t rows, int64_t cols, int64_t ld, int32_t b,
                                   int32_t nbr, int32_t nbc, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t blk = warp0; blk < static_cast<int64_t>(nbr) * nbc; blk += nwarps) {
    const int64_t r0 = (blk / nbc) * b, c0 = (blk % nbc) * b;
    int32_t s = 0;
    for (int e = lane; e < b * b; e += 32) {
      const int64_t r = r0 + e / b, c = c0 + e % b;
      if (r < rows && c < cols) { const int v = w[r * ld + c]; s += v < 0 ? -v : v; }
    }
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) out[blk] = s;
  }
}

__global__ void block_l2_f32_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int64_t ld, int32_t bh,
                                    int32_t bw, int32_t nbr, int32_t nbc, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t blk = warp0; blk < static_cast<int64_t>(nbr) * nbc; blk += nwarps) {
    const int64_t r0 = (blk / nbc) * bh, c0 = (blk % nbc) * bw;
    double s = 0.0;
    for (int e = lane; e < bh * bw; e += 32) {
      const int64_t r = r0 + e / bw, c = c0 + e % bw;
      if (r < rows && c < cols) { const double v = w[r * ld + c]; s += v * v; }
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[blk] = static_cast<float>(sqrt(s));
  }
}

// ------------------------------------------------------------------ packer: scan + compaction
// keep[nbr*nbc] -> per-row counts (one warp per block-row, ballot popcount)
__global__ void bsr_row_count_kernel(const uint8_t* __restrict__ keep, int32_t nbr, int32_t nbc,
                                     int32_t* __restrict__ row_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t br = warp0; br < nbr; br += nwarps) {
    int32_t cnt = 0;
    for (int32_t c = lane; c < nbc; c += 32) cnt += keep[br * nbc + c] ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) row_cnt[br] = cnt;
  }
}
// exclusive scan of row counts into row_ptr[0..nbr] (single CTA, chunked warp scans)
__global__ void bsr_row_scan_kernel(const int32_t* __restrict__ row_cnt, int32_t nbr, int32_t* __restrict__ row_ptr) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) { carry = 0; row_ptr[0] = 0; }
  __syncthreads();
  for (int32_t base = 0; base < nbr; base += blockDim.x) {
    const int32_t i = base + threadIdx.x;
    int32_t v = i < nbr ? row_cnt[i] : 0, incl = v;
    for (int o = 1; o < 32; o <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int32_t t = lane < nw ? warp_tot[lane] : 0, ti = t;
      for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += u; }
      warp_tot[lane] = ti - t;  // exclusive offsets per warp
      if (lane == 31) warp_tot[31] = ti - t;
    }
    __syncthreads();
    const int32_t total_prev = carry;
    if (i < nbr) row_ptr[i + 1] = total_prev + warp_tot[warp] + incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = total_prev + warp_tot[warp] + incl;
    __syncthreads();
  }
}
// slot[br*nbc+c] = output index of a kept block (row-major scan order == reference loop order), -1 = dropped
__global__ void bsr_slot_kernel(const uint8_t* __restrict__ keep, int32_t nbr, int32_t nbc,
                                const int32_t* __restrict__ row_ptr, int32_t* __restrict__ slot) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t br = warp0; br < nbr; br += nwarps) {
    int32_t base = row_ptr[br];
    for (int32_t c0 = 0; c0 < nbc; c0 += 32) {
      const int32_t c = c0 + lane;
      const bool k = c < nbc && keep[br * nbc + c];
      const uint32_t bal = __ballot_sync(0xffffffffu, k);
      if (c < nbc) slot[br * nbc + c] = k ? base + __popc(bal & ((1u << lane) - 1u)) : -1;
      base += __popc(bal);
    }
  }
}
// one warp per kept block: copy (zero padded) b x b bytes, record its column
__global__ void bsr_gather_i8_kernel(const int8_t* __restrict__ w, int64_t rows, int64_t cols, int64_t ld, int32_t b,
                                     const int32_t* __restrict__ slot, int32_t nbr, int32_t nbc,
                                     int32_t* __restrict__ col_idx, int8_t* __restrict__ blocks) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t blk = warp0; blk < static_cast<int64_t>(nbr) * nbc; blk += nwarps) {
    const int32_t s = slot[blk];
    if (s < 0) continue;
    const int64_t r0 = (blk / nbc) * b, c0 = (blk % nbc) * b;
    for (int e = lane; e < b * b; e += 32) {
      const int64_t r = r0 + e / b, c = c0 + e % b;
      blocks[static_cast<int64_t>(s) * b * b + e] = (r < rows && c < cols) ? w[r * ld + c] : static_cast<int8_t>(0);
    }
    if (lane == 0) col_idx[s] = static_cast<int32_t>(blk % nbc);
  }
}

__global__ void bsr_gather_f32_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int64_t ld, int32_t bh,
                                      int32_t bw, const int32_t* __restrict__ slot, int32_t nbr, int32_t nbc,
                                      int32_t* __restrict__ col_idx, float* __restrict__ blocks) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t blk = warp0; blk < static_cast<int64_t>(nbr) * nbc; blk += nwarps) {
    const int32_t s = slot[blk];
    if (s < 0) continue;
    const int64_t r0 = (blk / nbc) * bh, c0 = (blk % nbc) * bw;
    for (int e = lane; e < bh * bw; e += 32) {
      const int64_t r = r0 + e / bw, c = c0 + e % bw;
      blocks[static_cast<int64_t>(s) * bh * bw + e] = (r < rows && c < cols) ? w[r * ld + c] : 0.f;
    }
    if (lane == 0) col_idx[s] = static_cast<int32_t>(blk % nbc);
  }
}

// ------------------------------------------------------------------ quantiser (quantize.py:71-98)
__global__ void row_absmax_f32_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int64_t ld,
                                      float* __restrict__ absmax) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float m = 0.f;
    for (int64_t c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(w[r * ld + c]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) absmax[r] = m;
  }
}
__global__ void symmetric_scales_f32_kernel(const float* __restrict__ absmax, int64_t n, float* __restrict__ scales) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    scales[i] = fmaxf(__fdiv_rn(absmax[i], 127.0f), 1e-12f);
}
__global__ void quantize_rows_f32_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int64_t ld,
                                         const float* __restrict__ scales, int8_t* __restrict__ q) {
  const int64_t total = rows * cols;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / cols, c = idx - r * cols;
    const float v = rintf(__fdiv_rn(w[r * ld + c], scales[r]));   // np.rint(x / scale), float32
    q[idx] = static_cast<int8_t>(fminf(127.f, fmaxf(-128.f, v)));
  }
}

// ------------------------------------------------------------------ stand-alone epilogue pieces
__global__ void requant_i32_i8_kernel(const int32_t* __restrict__ acc, int8_t* __restrict__ out, int64_t n_outer,
                                      int64_t n_chan, int64_t n_inner, const float* __restrict__ sf,
                                      const int32_t* __restrict__ bias, int32_t relu,
                                      unsigned long long* __restrict__ sat_count) {
  const int64_t total = n_outer * n_chan * n_inner;
  uint32_t sat = 0;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = (idx / n_inner) % n_chan;
    int a = acc[idx];
    if (bias) a += bias[c];
    if (relu) a = max(a, 0);
    int q = __float2int_rn(__fmul_rn(__int2float_rn(a), sf[c]));
    if (q > 127) { q = 127; ++sat; } else if (q < -128) { q = -128; ++sat; }
    out[idx] = static_cast<int8_t>(q);
  }
  if (sat_count) {
    sat = __reduce_add_sync(0xffffffffu, sat);
    if ((threadIdx.x & 31) == 0 && sat) atomicAdd(sat_count, static_cast<unsigned long long>(sat));
  }
}
// relu_int8 / relu6_int8 / relu_int32 (golden_models.cpp:278-283, :323-330, :298-303), in place.  `hi` is the upper clamp:
// 127 for plain ReLU, int8(6.0f / scale) (computed on the host exactly as the reference does) for ReLU6.
__global__ void relu_i8_kernel(int8_t* __restrict__ d, int64_t n, int32_t hi) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int v = d[i];
    if (v < 0) v = 0;
    if (v > hi) v = hi;
    d[i] = static_cast<int8_t>(v);
  }
}
// Every second pixel of every second row of `planes` image planes: what a stride-2 / pad-0 1x1 convolution reads
// (conv2d_int8_im2col with stride 2, golden_models.cpp:883-933; the ResNet-50 downsample branches).  The row padding of
// the output is not touched (it stays zero).
__global__ void subsample2_i8_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int64_t planes, int32_t in_pitch,
                                     int32_t H, int32_t Ho, int32_t Wo, int32_t out_pitch) {
  const int64_t n = planes * Ho * Wo;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / Wo;
    const int xo = static_cast<int>(i - row * Wo);
    const int64_t pl = row / Ho;
    const int yo = static_cast<int>(row - pl * Ho);
    out[(pl * Ho + yo) * out_pitch + xo] = in[(pl * H + 2 * yo) * static_cast<int64_t>(in_pitch) + 2 * xo];
  }
}
// same, four output pixels per thread from one aligned 8-byte load (rows 8-byte aligned and long enough: checked on the host)
__global__ void subsample2_i8_vec_kernel(const int8_t* __restrict__ in, int8_t* __restrict__ out, int64_t planes, int32_t in_pitch,
                                         int32_t H, int32_t Ho, int32_t Wo, int32_t out_pitch) {
  const int q_per_row = (Wo + 3) >> 2;
  const int64_t n = planes * Ho * q_per_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / q_per_row;
    const int q = static_cast<int>(i - row * q_per_row);
    const int64_t pl = row / Ho;
    const int yo = static_cast<int>(row - pl * Ho);
    const uint2 v = *reinterpret_cast<const uint2*>(in + (pl * H + 2 * yo) * static_cast<int64_t>(in_pitch) + 8 * q);
    const uint32_t even = __byte_perm(v.x, v.y, 0x6420);
    int8_t* o = out + (pl * Ho + yo) * out_pitch + 4 * q;
    if (4 * q + 4 <= Wo) {
      *reinterpret_cast<uint32_t*>(o) = even;
    } else {
      for (int b = 0; 4 * q + b < Wo; ++b) o[b] = static_cast<int8_t>(even >> (8 * b));
    }
  }
}
__global__ void relu_i32_kernel(int32_t* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (d[i] < 0) d[i] = 0;
}

// ---- gemm_bsr_int8 (sw/golden/gemm_bsr_int8.py:16-104): the reference's float32 "golden" with per-row scales, replayed
// operation by operation.  Compatibility path: it is order-dependent float arithmetic on tiny shapes, so
// one thread owns one output element and walks (block-row, block, local row) in the reference's order.
// Step 1: re-quantise every stored block row with the scale of its global row (:74-79); T = float or double selects the
// dtype NumPy's promotion gives block / scale.
template <typename T>
__global__ void fp32compat_requant_kernel(const T* __restrict__ data, const int32_t* __restrict__ blk_row, int64_t nnz, int32_t bh,
                                          int32_t bw, const T* __restrict__ scales, int32_t n_scales, int8_t* __restrict__ q) {
  const int64_t total = nnz * bh * bw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = i / (bh * bw);
    const int r = static_cast<int>((i / bw) % bh);
    const int64_t g = static_cast<int64_t>(blk_row[b]) * bh + r;
    const T s = g < n_scales ? scales[g] : scales[0];
    T v = rint(data[i] / s);                       // IEEE divide, round half to even (np.rint)
    v = v < T(-128) ? T(-128) : (v > T(127) ? T(127) : v);
    q[i] = static_cast<int8_t>(v);
  }
}
// Step 2: C[m, n].  a64 / s64: scale_A / scales_B are float64 (NumPy then carries the products, and the += , in double and
// rounds to float32 on the store); otherwise every product and the add are float32.
__global__ void fp32compat_gemm_kernel(const int8_t* __restrict__ A, int64_t M, int64_t lda, const int32_t* __restrict__ indptr,
                                       const int32_t* __restrict__ indices, const int8_t* __restrict__ q, int32_t nbr, int32_t b,
                                       int64_t K, int64_t N, double scale_a, int32_t a64, const void* __restrict__ scales,
                                       int32_t s64, int32_t n_scales, float* __restrict__ C, int64_t ldc) {
  const int64_t total = M * N;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = idx / N, n = idx - m * N;
    float c = 0.f;
    for (int br = 0; br < nbr; ++br) {
      const int64_t rs = static_cast<int64_t>(br) * b;
      for (int i = indptr[br]; i < indptr[br + 1]; ++i) {
        const int64_t cs = static_cast<int64_t>(indices[i]) * b;
        if (n < cs || n >= cs + b) continue;
        const int j = static_cast<int>(n - cs);
        int t = 0;                                  // (A[:, rs:rs+b] @ Q^T)[m, j]
        for (int k = 0; k < b; ++k) t += static_cast<int>(A[m * lda + rs + k]) * static_cast<int>(q[(static_cast<int64_t>(i) * b + j) * b + k]);
        const float t32 = __int2float_rn(t);
        for (int r = 0; r < b; ++r) {
          if (rs + r >= K) break;
          const int64_t g = rs + r;
          const int64_t gi = g < n_scales ? g : 0;
          if (!a64 && !s64) {
            const float x = __fmul_rn(t32, static_cast<float>(scale_a));
            c = __fadd_rn(c, __fmul_rn(x, static_cast<const float*>(scales)[gi]));
          } else {
            const double x = a64 ? __dmul_rn(static_cast<double>(t32), scale_a)
                                 : static_cast<double>(__fmul_rn(t32, static_cast<float>(scale_a)));
            const double s = s64 ? static_cast<const double*>(scales)[gi] : static_cast<double>(static_cast<const float*>(scales)[gi]);
            c = static_cast<float>(__dadd_rn(static_cast<double>(c), __dmul_rn(x, s)));
          }
        }
      }
    }
    C[m * ldc + n] = c;
  }
}

__global__ void add_residual_i8_kernel(const int8_t* __restrict__ a, const int8_t* __restrict__ b,
                                       int8_t* __restrict__ out, int64_t n, float sa, float sb, float so) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = __fmul_rn(__int2float_rn(a[i]), sa), y = __fmul_rn(__int2float_rn(b[i]), sb);
    const int q = __float2int_rn(__fdiv_rn(__fadd_rn(x, y), so));
    out[i] = static_cast<int8_t>(min(127, max(-128, q)));
  }
}
// One thread per 4 adjacent outputs of a row: HBM-bound byte work, 4-byte stores, rows of the window re-used
// from registers.  Padding reads as -128 (maxpool2d_int8 initialises its maximum to -128, golden_models.cpp:549).
__global__ void maxpool_i8_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ out, int64_t n_planes, int32_t H,
                                  int32_t W, int32_t pool, int32_t stride, int32_t pad, int32_t Ho, int32_t Wo,
                                  int32_t in_pitch, int32_t out_pitch) {
  const int wq = (Wo + 3) >> 2;
  const int64_t total = n_planes * Ho * wq;
  const bool vec_out = (out_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx % wq), oh = static_cast<int>((idx / wq) % Ho);
    const int64_t pl = idx / (static_cast<int64_t>(wq) * Ho);
    const int ow0 = q * 4;
    int best[4] = {-128, -128, -128, -128};
    const int8_t* xp = x + pl * H * in_pitch;
    for (int ph = 0; ph < pool; ++ph) {
      const int ih = oh * stride + ph - pad;
      if (ih < 0 || ih >= H) continue;
      const int8_t* row = xp + static_cast<int64_t>(ih) * in_pitch;
      const int iw0 = ow0 * stride - pad;
      const int span = 3 * stride + pool;                 // input columns feeding the 4 outputs
      for (int j = 0; j < span; ++j) {
        const int iw = iw0 + j;
        if (iw < 0 || iw >= W) continue;
        const int v = row[iw];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int rel = j - o * stride;
          if (rel >= 0 && rel < pool) best[o] = max(best[o], v);
        }
      }
    }
    int8_t* orow = out + (pl * Ho + oh) * out_pitch + ow0;
    if (vec_out && ow0 + 3 < Wo) {
      const uint32_t pk = (static_cast<uint32_t>(best[0]) & 0xffu) | ((static_cast<uint32_t>(best[1]) & 0xffu) << 8) |
                          ((static_cast<uint32_t>(best[2]) & 0xffu) << 16) | ((static_cast<uint32_t>(best[3]) & 0xffu) << 24);
      *reinterpret_cast<uint32_t*>(orow) = pk;
    } else {
      for (int o = 0; o < 4 && ow0 + o < Wo; ++o) orow[o] = static_cast<int8_t>(best[o]);
    }
  }
}
// 3x3 / stride 2 / pad 1 (the ResNet stem pool) on 8-byte aligned rows: one thread = 4 adjacent outputs.  It
// reads the 16 input columns [8q-8, 8q+8) of three rows with 8-byte loads, reduces vertically with packed byte
// maxima (__vmaxs4), then horizontally on the even / odd / previous-odd columns.  HBM-bound byte work.
__global__ void maxpool3x3s2_i8_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ out, int64_t n_planes, int32_t H,
                                       int32_t W, int32_t Ho, int32_t Wo, int32_t in_pitch, int32_t out_pitch) {
  const int wq = (Wo + 3) >> 2;
  const int64_t total = n_planes * Ho * wq;
  constexpr uint32_t kNeg = 0x80808080u;   // -128 in every byte: what padding contributes to a maximum
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx % wq), oh = static_cast<int>((idx / wq) % Ho);
    const int64_t pl = idx / (static_cast<int64_t>(wq) * Ho);
    const int c0 = 8 * q;                              // first input column of the "current" 8 columns
    uint32_t v[4] = {kNeg, kNeg, kNeg, kNeg};          // columns c0-8 .. c0+7, vertical maximum
    const int8_t* xp = x + pl * static_cast<int64_t>(H) * in_pitch;
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int ih = 2 * oh - 1 + ph;
      if (ih < 0 || ih >= H) continue;
      const int8_t* row = xp + static_cast<int64_t>(ih) * in_pitch;
      uint2 lo = make_uint2(kNeg, kNeg), hi = make_uint2(kNeg, kNeg);
      if (c0 >= 8) lo = *reinterpret_cast<const uint2*>(row + c0 - 8);
      if (c0 < W) hi = *reinterpret_cast<const uint2*>(row + c0);
      uint32_t w4[4] = {lo.x, lo.y, hi.x, hi.y};
      if (c0 + 8 > W) {                                // columns past the row's end (pitch padding) are padding
#pragma unroll
        for (int j = 2; j < 4; ++j)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (c0 + (j - 2) * 4 + b >= W) w4[j] = (w4[j] & ~(0xffu << (8 * b))) | (0x80u << (8 * b));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __vmaxs4(v[j], w4[j]);
    }
    // output o (0..3) = max over columns c0 + 2o - 1, c0 + 2o, c0 + 2o + 1
    const uint32_t even = __byte_perm(v[2], v[3], 0x6420);        // columns c0+0, +2, +4, +6
    const uint32_t odd = __byte_perm(v[2], v[3], 0x7531);         // columns c0+1, +3, +5, +7
    const uint32_t podd = __byte_perm(v[1], odd, 0x6543);         // columns c0-1, +1, +3, +5
    const uint32_t r = __vmaxs4(__vmaxs4(even, odd), podd);
    int8_t* orow = out + (pl * Ho + oh) * static_cast<int64_t>(out_pitch) + 4 * q;
    if (4 * q + 3 < Wo) {
      *reinterpret_cast<uint32_t*>(orow) = r;
    } else {
      for (int o = 0; o < 4 && 4 * q + o < Wo; ++o) orow[o] = static_cast<int8_t>((r >> (8 * o)) & 0xffu);
    }
  }
}
// Same pooling, 16 outputs per thread: rows 16-byte aligned, six 16-byte loads (three input rows x 32 columns) and
// one 16-byte store per thread; the column left of the window comes from the previous 16 bytes' last byte.
__device__ __forceinline__ uint4 ldg_nc_128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__global__ void __launch_bounds__(256) maxpool3x3s2_i8_x16_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ out,
                                                                  uint32_t total, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                                                                  int32_t in_pitch, int32_t out_pitch, FastDiv d_wq, FastDiv d_ho) {
  constexpr uint32_t kNeg = 0x80808080u;
  // 32-bit index arithmetic with exact multiply-shift divisions: 64-bit / and % cost more than the pooling itself
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t t1 = fdiv(idx, d_wq);
    const int q = static_cast<int>(idx - t1 * d_wq.d);
    const uint32_t upl = fdiv(t1, d_ho);
    const int oh = static_cast<int>(t1 - upl * d_ho.d);
    const int64_t pl = upl;
    const int c0 = 32 * q;                             // first input column of this thread's 32
    const int8_t* xp = x + pl * static_cast<int64_t>(H) * in_pitch;
    uint4 a[3], b[3];
    uint32_t l[3];
#pragma unroll
    for (int ph = 0; ph < 3; ++ph) {
      const int ih = 2 * oh - 1 + ph;
      const bool ok = ih >= 0 && ih < H;
      const int8_t* row = xp + static_cast<int64_t>(ok ? ih : 0) * in_pitch + c0;
      a[ph] = make_uint4(kNeg, kNeg, kNeg, kNeg); b[ph] = a[ph]; l[ph] = kNeg;
      if (ok) {
        a[ph] = ldg_nc_128(row);                                   // c0 < W always (wq covers Wo)
        if (c0 + 16 < in_pitch) b[ph] = ldg_nc_128(row + 16);
        if (c0 > 0) l[ph] = *reinterpret_cast<const uint32_t*>(row - 4);
      }
    }
    // vertical maximum of the three rows (rows outside the image are all -128)
    uint32_t v[8];
    v[0] = __vmaxs4(__vmaxs4(a[0].x, a[1].x), a[2].x); v[1] = __vmaxs4(__vmaxs4(a[0].y, a[1].y), a[2].y);
    v[2] = __vmaxs4(__vmaxs4(a[0].z, a[1].z), a[2].z); v[3] = __vmaxs4(__vmaxs4(a[0].w, a[1].w), a[2].w);
    v[4] = __vmaxs4(__vmaxs4(b[0].x, b[1].x), b[2].x); v[5] = __vmaxs4(__vmaxs4(b[0].y, b[1].y), b[2].y);
    v[6] = __vmaxs4(__vmaxs4(b[0].z, b[1].z), b[2].z); v[7] = __vmaxs4(__vmaxs4(b[0].w, b[1].w), b[2].w);
    const uint32_t left = __vmaxs4(__vmaxs4(l[0], l[1]), l[2]);    // byte 3 = column c0 - 1
    if (c0 + 32 > W) {                                 // columns past the row's end (pitch padding) count as padding
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nv = W - (c0 + 4 * j);
        const uint32_t keep = nv >= 4 ? 0xffffffffu : (nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u));
        v[j] = (v[j] & keep) | (kNeg & ~keep);
      }
    }
    // output o (0..15) = max over columns c0 + 2o - 1, c0 + 2o, c0 + 2o + 1
    uint32_t r[4];
    uint32_t prev_odd_hi = left;                       // byte 3: column just left of the current 8
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t even = __byte_perm(v[2 * g], v[2 * g + 1], 0x6420);
      const uint32_t odd = __byte_perm(v[2 * g], v[2 * g + 1], 0x7531);
      const uint32_t podd = __byte_perm(prev_odd_hi, odd, 0x6543);   // columns -1, +1, +3, +5 of this group of 8
      r[g] = __vmaxs4(__vmaxs4(even, odd), podd);
      prev_odd_hi = v[2 * g + 1];
    }
    int8_t* orow = out + (pl * Ho + oh) * static_cast<int64_t>(out_pitch) + 16 * q;
    const int n_ok = Wo - 16 * q;
    if (n_ok >= 16) {
      *reinterpret_cast<uint4*>(orow) = make_uint4(r[0], r[1], r[2], r[3]);
    } else {
      int o = 0;
      if (n_ok >= 8) { *reinterpret_cast<uint2*>(orow) = make_uint2(r[0], r[1]); o = 8; }
      for (; o < n_ok; ++o) orow[o] = static_cast<int8_t>((r[o >> 2] >> (8 * (o & 3))) & 0xffu);
    }
  }
}
// eight lanes per plane (four planes per warp): (sum + HW/2) / HW with C truncating division (golden_models.cpp:619)
__global__ void avgpool_i8_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ out, int64_t n_planes,
                                  int32_t H, int32_t W, int32_t in_pitch) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int64_t grp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 3;
  const int64_t ngrp = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 3;
  const int hw = H * W;
  // rows that are 4-byte aligned words: one 32-bit load per 4 columns, packed byte sums by dp4a
  const bool words = (in_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0;
  const int wpr = (W + 3) >> 2;                        // words per row
  const int64_t n_iter = (n_planes + ngrp - 1) / ngrp;
  for (int64_t itn = 0; itn < n_iter; ++itn) {         // uniform trip count: the shuffles below need the whole warp
    const int64_t pl = grp0 + itn * ngrp;
    int s = 0;
    if (pl < n_planes) {
      if (words) {
        for (int i = sub; i < H * wpr; i += 8) {
          const int r = i / wpr, wi = i - r * wpr;
          uint32_t v = *reinterpret_cast<const uint32_t*>(x + (pl * H + r) * in_pitch + 4 * wi);
          const int nv = W - 4 * wi;
          if (nv < 4) v &= (1u << (8 * nv)) - 1u;
          s = __dp4a(static_cast<int>(v), 0x01010101, s);
        }
      } else {
        for (int i = sub; i < hw; i += 8) {
          const int r = i / W;
          s += x[(pl * H + r) * in_pitch + (i - r * W)];
        }
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (sub == 0 && pl < n_planes) out[pl] = static_cast<int8_t>(min(127, max(-128, (s + hw / 2) / hw)));
  }
}

// Rows of at most 16 pixels in 16-byte aligned, 16-byte-multiple pitches (the padded activation tensors): eight lanes per
// plane, lane `sub` sums rows sub, sub + 8, ... with one 16-byte load each - a warp reads four whole planes, contiguous
// when the pitch is 16.  Same rounding as above.
__global__ void __launch_bounds__(256) avgpool_i8_rows16_kernel(const int8_t* __restrict__ x, int8_t* __restrict__ out,
                                                                uint32_t n_planes, int32_t H, int32_t W, int32_t in_pitch) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t pl = t >> 3;
  const int sub = static_cast<int>(t & 7u);
  const int hw = H * W;
  uint32_t m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nv = W - 4 * i;
    m[i] = nv >= 4 ? 0xFFFFFFFFu : (nv <= 0 ? 0u : (1u << (8 * nv)) - 1u);
  }
  int s = 0;
  if (pl < n_planes) {
    const int8_t* base = x + static_cast<size_t>(pl) * H * in_pitch;
    for (int r = sub; r < H; r += 8) {
      const uint4 v = ldg_nc_128(base + static_cast<size_t>(r) * in_pitch);
      s = __dp4a(static_cast<int>(v.x & m[0]), 0x01010101, s);
      s = __dp4a(static_cast<int>(v.y & m[1]), 0x01010101, s);
      s = __dp4a(static_cast<int>(v.z & m[2]), 0x01010101, s);
      s = __dp4a(static_cast<int>(v.w & m[3]), 0x01010101, s);
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (sub == 0 && pl < n_planes) out[pl] = static_cast<int8_t>(min(127, max(-128, (s + hw / 2) / hw)));
}

}  // namespace accel
