// Block-sparse INT8 GEMM / implicit-im2col convolution on tcgen05 (sm_100a).
//
//   Y[m, br*14+h] = sum over stored blocks (br, bc) of  sum_w X[m, bc*14+w] * blk[h][w]
//   (sw/golden/golden_fc1_test.py:78-106), then the fused epilogue of SURVEY.md A.3.
//
// One CTA owns 128 activation rows x one group of block-rows (<= 32 -> 512 TMEM columns, one
// 16-column INT32 accumulator per block-row).  Every stored block is a true dense contraction:
// it is issued as (half of) a tcgen05.mma.kind::i8 with M=128 (activation rows), N=16 (the
// block-row, 14 channels + 2 zero rows), K=32 bytes (two 16-byte K slots = two adjacent K tiles).
//
// Warp roles (192 threads):
//   warps 0-3  activation producers: gather a [128 x 16 K-tiles] stage (GEMM rows or im2col
//              patches, reference K order) into the canonical K-major core-matrix layout,
//              re-striding 14 -> 16; afterwards the same four warps run the epilogue, one TMEM
//              lane (= activation row) per thread.
//   warp 4     MMA issuer (one elected lane) + TMEM allocation.
//   warp 5     weight loader: cp.async.bulk of pre-packed B-tile batches into a ring.
// Pipelines are mbarrier rings: x_full/x_empty (producers <-> MMA), w_full/w_empty (loader <-> MMA),
// acc_full (MMA -> epilogue).  tcgen05.commit releases the rings.
#pragma once
#include <cstdint>
#include <cstdio>

#include "../../include/accel_b200.h"
#include "plan.h"
#include "ptx.cuh"

namespace accel {

constexpr int kXStages = 2;
constexpr int kWStages = 3;
constexpr int kTileM = 128;
constexpr int kXTileStride = kTileM * kTile + 16;             // 2064: +16 B skews banks between K tiles
constexpr int kXStageBytes = kChunkTiles * kXTileStride;      // 33024
constexpr int kWStageBytes = kBatchBytes;                     // 8224
constexpr int kSmemX = 0;
constexpr int kSmemW = kSmemX + kXStages * kXStageBytes;
constexpr int kSmemBar = kSmemW + kWStages * ((kWStageBytes + 127) / 128 * 128);
constexpr int kSmemBytes = kSmemBar + 256;
constexpr int kThreads = 192;

struct TcParams {
  // activation source
  const int8_t* x;
  int64_t M;
  int32_t K;
  int64_t lda;
  int32_t x_align2;  // GEMM: base pointer and lda are even -> 16-bit loads
  // conv geometry (conv mode only)
  int32_t C, H, W, ksz, stride, pad, Ho, Wo;
  // plan
  const uint8_t* ws;
  const BatchInfo* batches;
  const GroupInfo* groups;
  int32_t n_groups;
  // epilogue
  accel_epilogue epi;
  void* out;
  accel_out_layout lay;
};

// ---------------------------------------------------------------------------------------------
// Repack: 14x14 reference blocks -> 16x32 B tiles (canonical K-major core matrices) + op meta.
// B tile bytes: [k_slot(2)][row(16)][16 B]; row n < 14 of slot s holds block row n (14 bytes).
__global__ void repack_blocks_kernel(const int8_t* __restrict__ blocks, uint8_t* __restrict__ ws,
                                     const OpSrc* __restrict__ op_src, const uint32_t* __restrict__ op_off,
                                     const uint32_t* __restrict__ op_meta_off, const uint16_t* __restrict__ op_meta,
                                     int64_t n_ops) {
  const int64_t op = blockIdx.x;
  if (op >= n_ops) return;
  const OpSrc s = op_src[op];
  uint8_t* dst = ws + op_off[op];
  for (int i = threadIdx.x; i < kBTileBytes; i += blockDim.x) {
    const int slot = i >> 8, row = (i >> 4) & 15, col = i & 15;
    const int32_t blk = slot ? s.blk_hi : s.blk_lo;
    int8_t v = 0;
    if (blk >= 0 && row < kBlock && col < kBlock) v = blocks[static_cast<int64_t>(blk) * 196 + row * kBlock + col];
    dst[i] = static_cast<uint8_t>(v);
  }
  if (threadIdx.x == 0) *reinterpret_cast<uint16_t*>(ws + op_meta_off[op]) = op_meta[op];
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int8_t sat8_count(int v, uint32_t& sat) {
  if (v > 127) { ++sat; return 127; }
  if (v < -128) { ++sat; return -128; }
  return static_cast<int8_t>(v);
}

template <bool kConv>
__global__ void __launch_bounds__(kThreads, 1) bsr_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* x_full = bars;                 // [kXStages]
  uint64_t* x_empty = bars + kXStages;     // [kXStages]
  uint64_t* w_full = bars + 2 * kXStages;  // [kWStages]
  uint64_t* w_empty = w_full + kWStages;   // [kWStages]
  uint64_t* acc_full = w_empty + kWStages; // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int gi = blockIdx.x % p.n_groups;
  const int64_t mtile = blockIdx.x / p.n_groups;
  const GroupInfo G = p.groups[gi];
  const int64_t m0 = mtile * kTileM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXStages; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc_dyn(tmem_slot, static_cast<uint32_t>(G.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =================================================================== activation producers
    const int tid = threadIdx.x;  // 0..127
    int step = 0;
    // conv: this thread's output position (row m0+tid) decomposed once
    int64_t img_base = 0;
    int ih0 = 0, iw0 = 0;
    bool row_ok = (m0 + tid) < p.M;
    if (kConv) {
      const int64_t m = m0 + tid;
      const int64_t P = static_cast<int64_t>(p.Ho) * p.Wo;
      const int64_t n = row_ok ? m / P : 0;
      const int pp = row_ok ? static_cast<int>(m - n * P) : 0;
      const int oh = pp / p.Wo, ow = pp - oh * p.Wo;
      ih0 = oh * p.stride - p.pad;
      iw0 = ow * p.stride - p.pad;
      img_base = n * static_cast<int64_t>(p.C) * p.H * p.W;
    }
    for (int b = G.batch_begin; b < G.batch_end; ++b) {
      const BatchInfo bi = p.batches[b];
      if (!(bi.flags & 1)) continue;  // one activation stage per K chunk
      const int s = step % kXStages;
      const uint32_t ph = (step / kXStages) & 1;
      mbar_wait(&x_empty[s], ph ^ 1);
      uint8_t* stage = smem + kSmemX + s * kXStageBytes;
      const int k_chunk0 = static_cast<int>(bi.chunk) * kChunkTiles * kBlock;
      if (!kConv) {
        // 16 threads cover one row's 16 K tiles (224 contiguous bytes); 8 rows per pass
        const int t = tid & 15;
        const int k0 = k_chunk0 + t * kBlock;
#pragma unroll 4
        for (int it = 0; it < kTileM / 8; ++it) {
          const int r = it * 8 + (tid >> 4);
          const int64_t m = m0 + r;
          uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
          if (m < p.M && k0 < p.K) {
            const int8_t* src = p.x + m * p.lda + k0;
            if (p.x_align2 && k0 + kBlock <= p.K) {
              const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
              const uint32_t h0 = s16[0], h1 = s16[1], h2 = s16[2], h3 = s16[3], h4 = s16[4], h5 = s16[5], h6 = s16[6];
              w0 = h0 | (h1 << 16); w1 = h2 | (h3 << 16); w2 = h4 | (h5 << 16); w3 = h6;
            } else {
              uint32_t bytes[16];
#pragma unroll
              for (int i = 0; i < 16; ++i)
                bytes[i] = (i < kBlock && k0 + i < p.K) ? static_cast<uint32_t>(static_cast<uint8_t>(src[i])) : 0u;
              w0 = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
              w1 = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | (bytes[7] << 24);
              w2 = bytes[8] | (bytes[9] << 8) | (bytes[10] << 16) | (bytes[11] << 24);
              w3 = bytes[12] | (bytes[13] << 8);
            }
          }
          *reinterpret_cast<uint4*>(stage + t * kXTileStride + r * kTile) = make_uint4(w0, w1, w2, w3);
        }
      } else {
        // one thread = one output position; walk k = (c, kh, kw) incrementally over the chunk
        int k = k_chunk0;
        const int kk = p.ksz * p.ksz;
        int c = k / kk;
        int rem = k - c * kk;
        int kh = rem / p.ksz, kw = rem - kh * p.ksz;
        const int8_t* img = p.x + img_base;
        for (int t = 0; t < kChunkTiles; ++t) {
          uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int i = 0; i < kBlock; ++i) {
            uint32_t v = 0;
            const int ih = ih0 + kh, iw = iw0 + kw;
            if (row_ok && k < p.K && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H) &&
                static_cast<unsigned>(iw) < static_cast<unsigned>(p.W))
              v = static_cast<uint8_t>(img[(static_cast<int64_t>(c) * p.H + ih) * p.W + iw]);
            w[i >> 2] |= v << ((i & 3) * 8);
            ++k;
            if (++kw == p.ksz) { kw = 0; if (++kh == p.ksz) { kh = 0; ++c; } }
          }
          *reinterpret_cast<uint4*>(stage + t * kXTileStride + tid * kTile) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&x_full[s]);
      ++step;
    }

    // =================================================================== epilogue (same 4 warps)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int64_t m = m0 + tid;
    const bool valid = m < p.M;
    const int flags = p.epi.flags;
    int64_t out_base = 0;
    if (valid) {
      const int64_t img = m / p.lay.rows_per_image;
      out_base = img * p.lay.image_stride + (m - img * p.lay.rows_per_image) * p.lay.row_stride;
    }
    uint32_t sat = 0;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int g = 0; g < G.n_rows; ++g) {
      uint32_t v[16];
      if ((G.nonempty >> g) & 1u) {
        tmem_ld16(tmem_base + lane_base + g * kTile, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      const int c0 = (G.br0 + g) * kBlock;
#pragma unroll
      for (int h = 0; h < kBlock; ++h) {
        const int c = c0 + h;
        const bool chan_ok = c < p.epi.n_channels;   // warp-uniform
        int acc = static_cast<int>(v[h]);
        if (chan_ok && p.epi.bias) acc += p.epi.bias[c];
        if (flags & ACCEL_RELU) acc = max(acc, 0);
        if (p.epi.chan_absmax && chan_ok) {
          const int a = valid ? (acc < 0 ? (acc == INT_MIN ? INT_MAX : -acc) : acc) : 0;
          const int wmax = __reduce_max_sync(0xffffffffu, a);
          if (lane == 0 && wmax > 0) atomicMax(p.epi.chan_absmax + c, wmax);
        }
        if (!valid || !chan_ok) continue;
        const int64_t o = out_base + static_cast<int64_t>(c) * p.lay.chan_stride;
        if (flags & ACCEL_OUT_I32) {
          reinterpret_cast<int32_t*>(p.out)[o] = acc;
        } else if (flags & ACCEL_OUT_F32) {
          reinterpret_cast<float*>(p.out)[o] = __fmul_rn(__int2float_rn(acc), p.epi.chan_scale[c]);
        } else {
          // golden_models.cpp:378-411, per channel: one float32 multiply, round-half-even, saturate
          const int q = __float2int_rn(__fmul_rn(__int2float_rn(acc), p.epi.chan_scale[c]));
          int8_t q8 = sat8_count(q, sat);
          if (p.epi.residual) {
            // golden_models.cpp:465-490: (main*s_main + res*s_res) / s_out, float32, no FMA
            const float a = __fmul_rn(__int2float_rn(q8), p.epi.res_scale_main);
            const float r = __fmul_rn(__int2float_rn(p.epi.residual[o]), p.epi.res_scale_res);
            const int z = __float2int_rn(__fdiv_rn(__fadd_rn(a, r), p.epi.res_scale_out));
            q8 = static_cast<int8_t>(min(127, max(-128, z)));
          }
          if ((flags & ACCEL_RELU_OUT) && q8 < 0) q8 = 0;   // relu_int8, golden_models.cpp:278-283
          reinterpret_cast<int8_t*>(p.out)[o] = q8;
        }
      }
    }
    if (p.epi.sat_count) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =================================================================== MMA issuer
    const uint32_t idesc = idesc_i8(kTileM, kTile);
    uint32_t inited = 0;
    int step = -1, wcount = 0;
    int xs = 0;
    for (int b = G.batch_begin; b < G.batch_end; ++b) {
      const BatchInfo bi = p.batches[b];
      if (bi.flags & 1) {
        ++step;
        xs = step % kXStages;
        mbar_wait(&x_full[xs], (step / kXStages) & 1);
      }
      const int ws_i = wcount % kWStages;
      mbar_wait(&w_full[ws_i], (wcount / kWStages) & 1);
      tc_fence_after();
      const uint8_t* wst = smem + kSmemW + ws_i * ((kWStageBytes + 127) / 128 * 128);
      const uint32_t x_addr = smem_u32(smem + kSmemX + xs * kXStageBytes);
      const uint32_t w_addr = smem_u32(wst);
      // Lane-parallel decode: lane i owns op i of the batch.  Everything that depends on the op's
      // metadata is computed once, in vector registers, for all ops at the same time; the issue loop
      // below only broadcasts one packed word per op (shfl -> warp-uniform -> uniform registers).
      const int n_ops = bi.n_ops;
      uint32_t mt = 0;
      if (lane < n_ops) mt = reinterpret_cast<const uint16_t*>(wst + n_ops * kBTileBytes)[lane];
      const uint32_t g = mt & 31u, win = mt >> 5;
      const uint32_t same = __match_any_sync(0xffffffffu, lane < n_ops ? g : 32u + lane);
      const uint32_t acc = ((inited >> g) & 1u) | ((same & ((1u << lane) - 1u)) != 0u ? 1u : 0u);
      // packed: bits 0..13 A start address >> 4, bits 14..22 D column, bit 23 accumulate
      const uint32_t packed = (((x_addr + win * kXTileStride) >> 4) & 0x3FFFu) | ((g * kTile) << 14) | (acc << 23);
      inited |= __reduce_or_sync(0xffffffffu, lane < n_ops ? (1u << g) : 0u);
      const uint64_t adesc_hi = smem_desc_kmajor(0, kXTileStride, 128);
      const uint64_t bdesc0 = smem_desc_kmajor(w_addr, 256, 128);
#pragma unroll
      for (int i = 0; i < kOpsPerBatch; ++i) {
        if (i < n_ops) {                                   // warp-uniform
          const uint32_t pk = __shfl_sync(0xffffffffu, packed, i);
          const uint64_t adesc = adesc_hi | static_cast<uint64_t>(pk & 0x3FFFu);
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((i * kBTileBytes) >> 4);
          if (elect_one()) mma_i8_ss(tmem_base + ((pk >> 14) & 0x1FFu), adesc, bdesc, idesc, (pk >> 23) & 1u);
        }
      }
      if (elect_one()) {
        mma_commit(&w_empty[ws_i]);
        if (bi.flags & 2) mma_commit(&x_empty[xs]);
        if (b == G.batch_end - 1) mma_commit(acc_full);
      }
      __syncwarp();
      ++wcount;
    }
    if (G.batch_begin == G.batch_end) {
      if (lane == 0) mbar_arrive(acc_full);  // nothing stored in this group: outputs are bias only
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // =================================================================== weight loader
    if (lane == 0) {
      int wcount = 0;
      for (int b = G.batch_begin; b < G.batch_end; ++b) {
        const BatchInfo bi = p.batches[b];
        const int ws_i = wcount % kWStages;
        mbar_wait(&w_empty[ws_i], ((wcount / kWStages) & 1) ^ 1);
        const uint32_t bytes = bi.n_ops * kBTileBytes + kBatchMetaBytes;
        mbar_arrive_expect_tx(&w_full[ws_i], bytes);
        bulk_g2s(smem + kSmemW + ws_i * ((kWStageBytes + 127) / 128 * 128), p.ws + static_cast<size_t>(bi.blob_off16) * 16,
                 bytes, &w_full[ws_i]);
        ++wcount;
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, static_cast<uint32_t>(G.tmem_cols));
  }
}

}  // namespace accel
