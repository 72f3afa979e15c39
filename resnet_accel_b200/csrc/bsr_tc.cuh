// Block-sparse INT8 GEMM / implicit-im2col convolution on tcgen05 (sm_100a) - activation-stationary design.
//
//   Y[m, br*14+h] = sum over stored blocks (br, bc) of  sum_w X[m, bc*14+w] * blk[h][w]
//   (sw/golden/golden_fc1_test.py:78-106), then the fused epilogue of SURVEY.md A.3.
//
// One CTA owns 128 activation rows (= 128 TMEM lanes) x one group of <= 11 block-rows.  Its 256 TMEM
// columns hold BOTH operands of the hot loop:
//     columns [0, 176)        one 16-column INT32 accumulator per block-row (zeroed up front)
//     columns [176, 248)      two activation stages of 9 K-tiles (16 bytes = 4 columns per tile, 14 data + 2 zero)
// so two CTAs share an SM and overlap each other's prologue / epilogue.  Every stored block is a true
// dense contraction, issued as (half of) one  tcgen05.mma.kind::i8  with A = activations FROM TMEM
// (M=128), B = the block-row's 16x32 weight tile from shared memory (N=16, K=32 = two adjacent K tiles).
// Measured on B200 (tools/probe): an SS-mode N=16 MMA is bound by the 4 KB shared-memory read of A
// (~45 cycles); with A in TMEM and several issuing warps the same MMA sustains ~16 cycles.
//
// Warp roles:
//   warps 0-7   activation producers, thread = activation row (TMEM lane), two halves of 4 warps that own
//               alternating stages.  GEMM: re-stride 14 -> 16 from the row-major activations.
//               Conv: implicit im2col in the reference K order (c, kh, kw) - the input rows a stage needs
//               are staged once in shared memory (coalesced, zero padded), then each thread walks its
//               patch with a compile-time (3x3) or incremental (general) pattern and writes 16-byte
//               tiles straight into TMEM with tcgen05.st.  Afterwards the same warps run the epilogue.
//   warps 8..   MMA issuers (kIssuers warps, ops dealt round-robin) + TMEM allocation.
//   last warp   weight loader: cp.async.bulk of pre-packed B-tile batches into a ring.
#pragma once
#include <climits>
#include <cstdint>
#include <cstdio>

#include "../../include/accel_b200.h"
#include "plan.h"
#include "ptx.cuh"

namespace accel {

constexpr int kTileM = 128;
constexpr int kProducerWarps = 8;
constexpr int kIssuers = 4;
constexpr int kThreads = (kProducerWarps + kIssuers + 1) * 32;
constexpr int kXStages = 2;
constexpr int kWStages = 3;
constexpr int kAccCols = kMaxGroupRows * kTile;                 // 176
constexpr int kXStageCols = kChunkTiles * 4;                    // 36
constexpr int kTmemCols = 256;
static_assert(kAccCols + kXStages * kXStageCols <= kTmemCols, "TMEM budget");
constexpr int kWStageBytes = (kBatchBytes + 127) / 128 * 128;
constexpr int kSmemW = 0;
constexpr int kSmemBar = kSmemW + kWStages * kWStageBytes;
constexpr int kSmemTab = kSmemBar + 128;                        // per-halo-row (image, first input row) tables
constexpr int kMaxHaloRows = 40;
constexpr int kSmemHalo = kSmemBar + 128 + 2 * kMaxHaloRows * 4 + 64;   // two halo buffers follow (conv only)
constexpr int kGemmStagePitch = 148;                            // 37 words: odd pitch -> conflict-free per-row reads
constexpr int kGemmStageBytes = kTileM * kGemmStagePitch + 64;  // one staging buffer per producer half (GEMM only)

struct TcParams {
  // activation source
  const int8_t* x;
  int64_t M;
  int32_t K;
  int64_t lda;
  int32_t x_align2;  // GEMM: base pointer and lda are even -> 16-bit loads
  int32_t gemm_staged;  // GEMM: base pointer and lda are multiples of 16 -> coalesced 16-byte loads via shared memory
  // conv geometry (conv mode only)
  int32_t C, H, W, ksz, stride, pad, Ho, Wo;
  int32_t halo_pitch, halo_rows, halo_bytes;   // per producer half: [14 ch][halo_rows][ksz][halo_pitch]
  int32_t halo_lpr, halo_vec;                  // lanes per halo row (pow2 >= words per row); rows are 4-byte copyable
  uint32_t halo_perch_magic;                   // ceil(2^32 / (halo_rows*3)): row / per_ch by __umulhi
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 3x3 fast path: one stage = 9 K tiles = 126 k = exactly 14 input channels x 9 taps, so the (channel, tap) ->
// (tile, byte) map is a compile-time pattern.  `rowp` points at this thread's patch origin inside the halo.
__device__ __forceinline__ void gather3x3_to_tmem(const uint8_t* rowp, int chan_stride, int pitch, uint32_t tcol) {
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < 14; ++c) {
    const uint8_t* pc = rowp + c * chan_stride;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const uint8_t* pr = pc + kh * pitch;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int kl = c * 9 + kh * 3 + kw;        // compile-time
        const int i = kl % 14;
        const uint32_t b = pr[kw];
        if ((i & 3) == 0) w[i >> 2] = b; else w[i >> 2] |= b << ((i & 3) * 8);
        if (i == 13) {
          tmem_st4(tcol + (kl / 14) * 4, w[0], w[1], w[2], w[3]);
          w[3] = 0u;
        }
      }
    }
  }
}

template <bool kConv>
__global__ void __launch_bounds__(kThreads, 2) bsr_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* x_full = bars;                 // [2]  count 128 (one producer half)
  uint64_t* x_empty = bars + 2;            // [2]  count kIssuers
  uint64_t* w_full = bars + 4;             // [kWStages] count 1 (+tx bytes)
  uint64_t* w_empty = w_full + kWStages;   // [kWStages] count kIssuers
  uint64_t* acc_full = w_empty + kWStages; // [1]  count kIssuers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int gi = blockIdx.x % p.n_groups;
  const int64_t mtile = blockIdx.x / p.n_groups;
  const GroupInfo G = p.groups[gi];
  const int64_t m0 = mtile * kTileM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_empty[s], kIssuers); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kIssuers); }
    mbar_init(acc_full, kIssuers);
    fence_mbar_init();
  }
  if (warp == kProducerWarps) {
    tmem_alloc_dyn(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  int* tab_n = reinterpret_cast<int*>(smem + kSmemTab);        // image index of halo row rr (-1: past the batch)
  int* tab_ih0 = tab_n + kMaxHaloRows;                          // first input row (oh*stride - pad) of halo row rr
  if (kConv) {  // halo pads (left/right columns, rows never written) must read as zero
    uint32_t* h = reinterpret_cast<uint32_t*>(smem + kSmemHalo);
    for (int i = threadIdx.x; i < (2 * p.halo_bytes) / 4; i += kThreads) h[i] = 0u;
    if (threadIdx.x < p.halo_rows) {
      const int Rg = static_cast<int>(m0 / p.Wo) + threadIdx.x;
      const int n = Rg / p.Ho;
      tab_n[threadIdx.x] = (static_cast<int64_t>(n) * p.Ho * p.Wo < p.M) ? n : -1;
      tab_ih0[threadIdx.x] = (Rg - n * p.Ho) * p.stride - p.pad;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {  // zero the accumulators: every MMA accumulates, so issue order across warps is free
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (int c = 0; c < G.n_rows * kTile; c += 4) tmem_st4(tmem_base + lane_base + c, 0u, 0u, 0u, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kProducerWarps) {
    // =================================================================== activation producers
    const int half = warp >> 2;                 // owns stages with (step & 1) == half
    const int tid = threadIdx.x & 127;          // activation row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t xcol = tmem_base + lane_base + kAccCols + half * kXStageCols;
    const int64_t m = m0 + tid;
    const bool row_ok = m < p.M;
    // conv: this thread's output position and the tile's first output row
    int64_t R0 = 0;
    int rl = 0, ow = 0, oh = 0;
    int64_t img = 0;
    if (kConv) {
      const int64_t P = static_cast<int64_t>(p.Ho) * p.Wo;
      R0 = m0 / p.Wo;                                        // global output-row id (n*Ho + oh) of the tile start
      const int64_t mm = row_ok ? m : m0;
      const int64_t R = mm / p.Wo;
      ow = static_cast<int>(mm - R * p.Wo);
      rl = static_cast<int>(R - R0);
      img = mm / P;
      oh = static_cast<int>(R - img * p.Ho);
    }
    uint8_t* halo = smem + kSmemHalo + half * p.halo_bytes;
    int step = 0;
    for (int b = G.batch_begin; b < G.batch_end; ++b) {
      const BatchInfo bi = p.batches[b];
      if (!(bi.flags & 1)) continue;            // one activation stage per K chunk
      const int my = (step & 1) == half;
      const int use = step >> 1;
      ++step;
      if (!my) continue;
      mbar_wait(&x_empty[half], (use & 1) ^ 1);
      tc_fence_after();
      const int k_chunk0 = static_cast<int>(bi.chunk) * kChunkTiles * kBlock;
      if (!kConv) {
        // ---------------- GEMM rows: 9 tiles x 14 bytes, re-strided to 16.  Control flow is warp-uniform
        // (tcgen05.st is .sync.aligned): rows past M load nothing but still store their zeros.
        if (p.gemm_staged) {
          // Coalesced path (16-byte aligned rows): the 128 x 144-byte window that covers this stage is pulled
          // with 16-byte loads (9 consecutive lanes per row), parked in shared memory with an odd word pitch,
          // then every thread re-strides its own row conflict-free.
          uint8_t* stg = smem + kSmemHalo + half * kGemmStageBytes;
          const int a0 = k_chunk0 & ~15;
          const int shift = k_chunk0 - a0;                    // even, <= 14
          uint4 v[kChunkTiles];
#pragma unroll
          for (int it = 0; it < kChunkTiles; ++it) {
            const int item = it * 128 + tid;
            const int row = item / kChunkTiles, ch = item - row * kChunkTiles;
            const int64_t gm = m0 + row;
            const int ka = a0 + ch * 16;
            v[it] = make_uint4(0u, 0u, 0u, 0u);
            if (gm < p.M && ka < p.K) {
              const int8_t* g = p.x + gm * p.lda + ka;
              if (ka + 16 <= p.K) {
                v[it] = *reinterpret_cast<const uint4*>(g);
              } else {                                        // the one chunk per row that straddles K
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                for (int j = 0; j < 16; ++j)
                  if (ka + j < p.K) w[j >> 2] |= static_cast<uint32_t>(static_cast<uint8_t>(g[j])) << ((j & 3) * 8);
                v[it] = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
#pragma unroll
          for (int it = 0; it < kChunkTiles; ++it) {
            const int item = it * 128 + tid;
            const int row = item / kChunkTiles, ch = item - row * kChunkTiles;
            uint32_t* d = reinterpret_cast<uint32_t*>(stg + row * kGemmStagePitch + ch * 16);
            d[0] = v[it].x; d[1] = v[it].y; d[2] = v[it].z; d[3] = v[it].w;
          }
          named_bar_sync(1 + half, 128);
          const uint16_t* s16 = reinterpret_cast<const uint16_t*>(stg + tid * kGemmStagePitch + shift);
#pragma unroll
          for (int t = 0; t < kChunkTiles; ++t) {
            uint32_t h[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) h[j] = s16[t * 7 + j];
            tmem_st4(xcol + t * 4, h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6]);
          }
          named_bar_sync(1 + half, 128);
        } else {
          const int8_t* src = p.x + m * p.lda + k_chunk0;
          if (p.x_align2 && k_chunk0 + kChunkTiles * kBlock <= p.K) {
            const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src);
#pragma unroll
            for (int t = 0; t < kChunkTiles; ++t) {
              uint32_t h[7];
#pragma unroll
              for (int j = 0; j < 7; ++j) h[j] = row_ok ? static_cast<uint32_t>(s16[t * 7 + j]) : 0u;
              tmem_st4(xcol + t * 4, h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6]);
            }
          } else {
#pragma unroll 1
            for (int t = 0; t < kChunkTiles; ++t) {
              uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
              for (int i = 0; i < kBlock; ++i) {
                const int k = k_chunk0 + t * kBlock + i;
                const uint32_t v = (row_ok && k < p.K) ? static_cast<uint32_t>(static_cast<uint8_t>(src[t * kBlock + i])) : 0u;
                w[i >> 2] |= v << ((i & 3) * 8);
              }
              tmem_st4(xcol + t * 4, w[0], w[1], w[2], w[3]);
            }
          }
        }
      } else if (p.halo_bytes > 0) {
        // ---------------- 3x3 conv: stage the 14 channels' input rows, then the compile-time gather
        const int c0 = static_cast<int>(bi.chunk) * 14;
        const int per_ch = p.halo_rows * 3;
        const int wpr = (p.W + 3) >> 2;                       // 4-byte words per input row
        const int total_rows = 14 * per_ch;
        const int lpr = p.halo_lpr;                           // lanes per halo row (power of two >= wpr)
        const int wl = tid & (lpr - 1);
        const int rpp = 128 / lpr;                            // halo rows per pass
        if (wl < wpr) {
          for (int row = tid / lpr; row < total_rows; row += rpp) {
            const int cl = __umulhi(static_cast<unsigned>(row), p.halo_perch_magic);   // row / per_ch
            const int idx = row - cl * per_ch;
            const int rr = idx / 3, kh = idx - rr * 3;
            const int n = tab_n[rr];
            const int ih = tab_ih0[rr] + kh;
            const int c = c0 + cl;
            const bool ok = c < p.C && n >= 0 && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H);
            uint8_t* d = halo + row * p.halo_pitch + 4 + wl * 4;
            if (p.halo_vec) {
              // asynchronous 4-byte copies (LDGSTS), zero-filled for padding rows: nothing waits until the end
              const int8_t* g = ok ? p.x + ((static_cast<int64_t>(n) * p.C + c) * p.H + ih) * p.W + wl * 4 : p.x;
              cp_async4_zfill(d, g, ok);
            } else {
              uint32_t v = 0u;
              if (ok) {
                const int8_t* g = p.x + ((static_cast<int64_t>(n) * p.C + c) * p.H + ih) * p.W + wl * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (wl * 4 + j < p.W) v |= static_cast<uint32_t>(static_cast<uint8_t>(g[j])) << (8 * j);
              }
              *reinterpret_cast<uint32_t*>(d) = v;
            }
          }
        }
        cp_async_wait_all();
        named_bar_sync(1 + half, 128);
        const int chan_stride = p.halo_rows * 3 * p.halo_pitch;
        const uint8_t* rowp = halo + rl * 3 * p.halo_pitch + ow * p.stride + 4 - p.pad;
        gather3x3_to_tmem(rowp, chan_stride, p.halo_pitch, xcol);
        named_bar_sync(1 + half, 128);                        // halo is reused by this half's next stage
      } else {
        // ---------------- general k x k (1x1 downsample, 7x7 stem): incremental (c, kh, kw) walk from global
        int k = k_chunk0;
        const int kk = p.ksz * p.ksz;
        int c = k / kk;
        int rem = k - c * kk;
        int kh = rem / p.ksz, kw = rem - kh * p.ksz;
        const int8_t* im = p.x + img * static_cast<int64_t>(p.C) * p.H * p.W;
        const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
#pragma unroll 1
        for (int t = 0; t < kChunkTiles; ++t) {
          uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int i = 0; i < kBlock; ++i) {
            uint32_t v = 0;
            const int ih = ih0 + kh, iw = iw0 + kw;
            if (row_ok && k < p.K && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H) &&
                static_cast<unsigned>(iw) < static_cast<unsigned>(p.W))
              v = static_cast<uint8_t>(im[(static_cast<int64_t>(c) * p.H + ih) * p.W + iw]);
            w[i >> 2] |= v << ((i & 3) * 8);
            ++k;
            if (++kw == p.ksz) { kw = 0; if (++kh == p.ksz) { kh = 0; ++c; } }
          }
          tmem_st4(xcol + t * 4, w[0], w[1], w[2], w[3]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&x_full[half]);
    }

    // =================================================================== epilogue (same 8 warps)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool valid = row_ok;
    const int flags = p.epi.flags;
    int64_t out_base = 0;
    if (valid) {
      const int64_t im = m / p.lay.rows_per_image;
      out_base = im * p.lay.image_stride + (m - im * p.lay.rows_per_image) * p.lay.row_stride;
    }
    uint32_t sat = 0;
    for (int g = half; g < G.n_rows; g += 2) {
      uint32_t v[16];
      tmem_ld16(tmem_base + lane_base + g * kTile, v);
      tmem_ld_wait();
      const int cb = (G.br0 + g) * kBlock;
#pragma unroll
      for (int h = 0; h < kBlock; ++h) {
        const int c = cb + h;
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
  } else if (warp < kProducerWarps + kIssuers) {
    // =================================================================== MMA issuers
    const int me = warp - kProducerWarps;
    const uint32_t idesc = idesc_i8(kTileM, kTile);
    int step = -1, wcount = 0, xs = 0;
    for (int b = G.batch_begin; b < G.batch_end; ++b) {
      const BatchInfo bi = p.batches[b];
      if (bi.flags & 1) {
        ++step;
        xs = step & 1;
        mbar_wait(&x_full[xs], (step >> 1) & 1);
      }
      const int ws_i = wcount % kWStages;
      mbar_wait(&w_full[ws_i], (wcount / kWStages) & 1);
      tc_fence_after();
      const uint8_t* wst = smem + kSmemW + ws_i * kWStageBytes;
      const uint32_t w_addr = smem_u32(wst);
      // Lane-parallel decode: lane i owns op i of the batch; the issue loop only broadcasts one packed
      // word per op (shfl -> warp-uniform -> uniform registers for the descriptors).
      const int n_ops = bi.n_ops;
      uint32_t mt = 0;
      if (lane < n_ops) mt = reinterpret_cast<const uint16_t*>(wst + n_ops * kBTileBytes)[lane];
      const uint32_t g = mt & 15u, win = mt >> 4;
      // packed: bits 0..8 A column, bits 9..17 D column
      const uint32_t packed = (kAccCols + xs * kXStageCols + win * 4) | ((g * kTile) << 9);
      const uint64_t bdesc0 = smem_desc_kmajor(w_addr, 256, 128);
#pragma unroll
      for (int i = 0; i < kOpsPerBatch / kIssuers; ++i) {
        const int op = i * kIssuers + me;                  // ops are dealt round-robin to the issuing warps
        if (op < n_ops) {                                  // warp-uniform
          const uint32_t pk = __shfl_sync(0xffffffffu, packed, op);
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((op * kBTileBytes) >> 4);
          if (elect_one()) mma_i8_ts(tmem_base + (pk >> 9), tmem_base + (pk & 0x1FFu), bdesc, idesc, 1u);
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
        bulk_g2s(smem + kSmemW + ws_i * kWStageBytes, p.ws + static_cast<size_t>(bi.blob_off16) * 16, bytes,
                 &w_full[ws_i]);
        ++wcount;
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, kTmemCols);
  }
}

}  // namespace accel
