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
// (M=128), B = 16x32 weight tiles from shared memory (K=32 = two adjacent K tiles); block-rows that are
// adjacent and use the same K window are issued together as one MMA with N = 16*len.
//
// Warp roles (14 warps):
//   warps 0-7   activation producers, thread = activation row (TMEM lane), two halves of 4 warps that own
//               alternating stages; they re-stride 14 -> 16 and write tiles into TMEM with tcgen05.st.
//                 GEMM    rows come from a shared-memory ring filled by the activation loader (16-byte
//                         cp.async), re-strided with compile-time byte permutes.
//                 CONV<KS> implicit im2col in the reference K order (c, kh, kw): the input rows a stage needs
//                         sit in a shared-memory ring (each input row once, zero padded); every (c, kh)
//                         contributes KS contiguous bytes, merged with compile-time byte permutes.
//                 DIRECT  any kernel size / unaligned tensors: per-element gather from global memory.
//               Afterwards the same warps run the fused epilogue (per-channel constants cached in smem).
//   warps 8-11  MMA issuers.  The whole issue loop runs on one elected thread in the uniform datapath:
//               per MMA one 8-byte record from the kernel parameter bank, a handful of uniform ALU ops, UTCIMMA.
//   warp 12     weight loader: cp.async.bulk of pre-packed B-tile batches into a 3-stage ring.
//   warp 13     activation loader: cp.async (LDGSTS) of GEMM rows / conv input rows into the activation ring,
//               running up to 3 stages ahead of the producers.
#pragma once
#include <cuda.h>

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
constexpr int kWarpWLoad = kProducerWarps + kIssuers;
constexpr int kWarpALoad = kWarpWLoad + 1;
constexpr int kThreads = (kWarpALoad + 1) * 32;                 // 448
constexpr int kWStages = 3;
constexpr int kWStageBytes = kTilesPerBatch * kBTileBytes;      // 16 KB
constexpr int kAccCols = kMaxGroupRows * kTile;                 // 176
constexpr int kXStageCols = kChunkTiles * 4;                    // 36
constexpr int kTmemCols = 256;
static_assert(kAccCols + 2 * kXStageCols <= kTmemCols, "TMEM budget");
constexpr int kMaxRingSlots = 3;
constexpr int kMaxHaloRows = 48;                                // input rows held per channel plane
constexpr int kMaxOutRows = 40;                                 // output rows touched by one 128-row tile
constexpr int kGemmSlotBytes = kTileM * 144;                    // 128 rows x 9 x 16 B
// shared-memory map
constexpr int kSmemW = 0;
constexpr int kSmemBar = kSmemW + kWStages * kWStageBytes;      // barriers (256 B)
constexpr int kSmemScale = kSmemBar + 256;                      // float [176]
constexpr int kSmemBias = kSmemScale + kAccCols * 4;            // int32 [176]
constexpr int kSmemRowOff = kSmemBias + kAccCols * 4;           // int64 [kMaxHaloRows]
constexpr int kSmemHBase = kSmemRowOff + kMaxHaloRows * 8;      // int32 [kMaxOutRows] halo row of out-row's first input row
constexpr int kSmemOutRow = kSmemHBase + kMaxOutRows * 4;       // int32 [kMaxOutRows] image | flags
constexpr int kSmemRing = (kSmemOutRow + kMaxOutRows * 4 + 127) / 128 * 128;
constexpr int kRingSlack = 32;                                  // the gathers read whole words past a row's last byte

enum TcMode { kModeGemm = 0, kModeConv3 = 1, kModeConv7 = 2, kModeDirect = 3 };

// Exact unsigned 32-bit division by a run-time constant (Granlund-Montgomery, the "round-up" variant): the host
// computes (mul, shr) once, the device pays a multiply-high, a subtract and two shifts instead of ~25 instructions.
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d ? d : 1u, 0u, 0u};
  if (f.d > 1) {
    uint32_t l = 0;
    while ((1ull << l) < f.d) ++l;                           // l = ceil(log2 d)
    f.mul = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1);
    f.shr = l - 1;
  }
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  if (f.d == 1) return n;
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> 1)) >> f.shr;
}

struct TcParams {
  // activation source
  const int8_t* x;
  int64_t M;
  int32_t K;
  int64_t lda;
  int32_t x_align2;     // DIRECT GEMM: base pointer and lda are even -> 16-bit loads
  int32_t x_small;      // the activation tensor spans < 4 GiB: 32-bit byte offsets
  int32_t conv;         // DIRECT: 0 = GEMM rows, 1 = convolution gather
  // conv geometry
  int32_t C, H, W, ksz, stride, pad, Ho, Wo;
  int32_t Wp;           // bytes between input rows (>= W); planes and images are dense in that pitch
  // activation ring
  int32_t ring_slots;   // 2 or 3
  int32_t slot_bytes;   // bytes per ring slot (16-byte multiple)
  int32_t halo_rows;    // allocated input rows per channel plane
  int32_t halo_pitch;   // bytes per staged input row: halo_lpad zero bytes, W data bytes, >= pad zero bytes
  int32_t halo_lpad;    // 4 or 16
  int32_t halo_nch;     // channel planes per stage
  int32_t halo_vec;     // copy granularity: 16 / 4 (cp.async of aligned rows) or 0 (any alignment, through registers)
  int32_t halo_ipr;     // copy items per input row
  uint32_t ipr_magic;   // ceil(2^32 / halo_ipr)
  int32_t use_tma;      // activations arrive by TMA tensor tiles (16-byte aligned rows): GEMM 144 B x 128 rows per stage,
                        // conv one [planes][halo_rows][halo_pitch] box per image the tile touches (out-of-bounds = 0)
  int32_t seg_bytes;    // conv + TMA: bytes reserved for one image segment inside a ring slot (128-byte multiple)
  int32_t box_bytes;    // conv + TMA: bytes one tensor tile delivers (planes x rows x pitch)
  // weights
  const uint8_t* blob;
  // epilogue
  accel_epilogue epi;
  int32_t res_fast;     // residual divide may use the 3-instruction exact sequence (checked on the host)
  float res_rcp;        // RN(1 / res_scale_out)
  void* out;
  accel_out_layout lay;
  int32_t out_small;    // every output / residual element offset fits 31 bits
  FastDiv d_wo, d_ho, d_groups, d_rpi, d_rowlen;   // divisions by Wo, Ho, groups of this launch, rows_per_image, row_len
  long long* timeline;  // developer aid: per-CTA clock64 stamps (32 per CTA) when non-null
  int32_t dbg_flags;    // developer aid: bit 0 = skip the epilogue's global stores
};

struct TcLaunch {
  alignas(64) CUtensorMap tmap;   // activation tensor (use_tma)
  TcParams p;
  uint32_t n_groups;
  uint32_t pad_;
  GroupRec groups[kMaxGroupsL];
  uint32_t batches[kMaxBatchesL];
  OpRec ops[kMaxOpsL];
};
static_assert(sizeof(TcLaunch) <= 32764, "kernel parameter block too large");

// ---------------------------------------------------------------------------------------------
// Repack: 14x14 reference blocks -> 16x32 B tiles.  B tile bytes: [row_group(2)][k_slot(2)][8 rows][16 B]
// (K-major core matrices: LBO = 128 between the K halves, SBO = 256 between 8-row groups), so `len`
// consecutive tiles form one N = 16*len operand.  Row n < 14 of slot s holds block row n (14 bytes).
__global__ void repack_blocks_kernel(const int8_t* __restrict__ blocks, uint8_t* __restrict__ blob,
                                     const TileSrc* __restrict__ src, int64_t n_tiles) {
  const int64_t t = blockIdx.x;
  if (t >= n_tiles) return;
  const TileSrc s = src[t];
  uint8_t* dst = blob + t * kBTileBytes;
  for (int i = threadIdx.x; i < kBTileBytes; i += blockDim.x) {
    const int rg = i >> 8, slot = (i >> 7) & 1, r8 = (i >> 4) & 7, col = i & 15;
    const int row = rg * 8 + r8;
    const int32_t blk = slot ? s.blk_hi : s.blk_lo;
    int8_t v = 0;
    if (blk >= 0 && row < kBlock && col < kBlock) v = blocks[static_cast<int64_t>(blk) * 196 + row * kBlock + col];
    dst[i] = static_cast<uint8_t>(v);
  }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ int8_t cvt_sat_s8(float f) {   // round-half-even + saturate in one instruction
  int v;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(v) : "f"(f));
  return static_cast<int8_t>(v);
}

// ---- 14-in-16 tile stream: insert the `nb` bytes [src_byte, src_byte + nb) of `v` at stream byte `pos`.
// Stream byte d lives in tile d / 14 at byte d % 14; a tile is 4 words, its bytes 14, 15 stay zero.
template <int pos, int nb, int src_byte = 0>
__device__ __forceinline__ void stream_insert(uint32_t (&w)[36], uint32_t v) {
  if constexpr (nb > 0) {
    constexpr int t = pos / 14, b = pos % 14;
    constexpr int word = t * 4 + b / 4, byte = b % 4;
    constexpr int room_w = 4 - byte, room_t = 14 - b;
    constexpr int n = nb < room_w ? (nb < room_t ? nb : room_t) : (room_w < room_t ? room_w : room_t);
    constexpr uint32_t sel = ((byte <= 0 && 0 < byte + n) ? (4u + src_byte + 0 - byte) : 0u) |
                             (((byte <= 1 && 1 < byte + n) ? (4u + src_byte + 1 - byte) : 1u) << 4) |
                             (((byte <= 2 && 2 < byte + n) ? (4u + src_byte + 2 - byte) : 2u) << 8) |
                             (((byte <= 3 && 3 < byte + n) ? (4u + src_byte + 3 - byte) : 3u) << 12);
    w[word] = __byte_perm(w[word], v, sel);
    stream_insert<pos + n, nb - n, src_byte + n>(w, v);
  }
}
template <int t0, int t1>
__device__ __forceinline__ void stream_flush(const uint32_t (&w)[36], uint32_t xcol) {
  if constexpr (t0 < t1) {
    tmem_st4(xcol + t0 * 4, w[t0 * 4], w[t0 * 4 + 1], w[t0 * 4 + 2], w[t0 * 4 + 3]);
    stream_flush<t0 + 1, t1>(w, xcol);
  }
}

// ---- CONV<KS>: one stage = 126 stream bytes = 126 / KS groups; group (c, kh) = KS contiguous bytes of one
// staged input row.  `abase` = shared address of the aligned word that holds this thread's first tap of
// (channel plane 0, this thread's first input row); `sh8` = 8 * (byte offset inside that word).
template <int KS, int j, int e, int NG>
__device__ __forceinline__ void conv_extract(uint32_t (&w)[36], const uint32_t (&g)[NG], uint32_t sh8) {
  constexpr int NE = (KS + 3) / 4;          // extracted registers per group
  if constexpr (e < NE) {
    const uint32_t v = __funnelshift_r(g[e], g[e + 1], sh8);
    stream_insert<j * KS + 4 * e, (KS - 4 * e) < 4 ? (KS - 4 * e) : 4>(w, v);
    conv_extract<KS, j, e + 1, NG>(w, g, sh8);
  }
}
// Groups are processed in batches of kGB: all shared-memory loads of a batch are issued before the first
// permute that consumes them, so the ~30-cycle load latency is paid once per batch instead of once per group.
template <int KS>
struct ConvBatch {
  static constexpr int GPS = 126 / KS;
  static constexpr int NW = (KS + 3 + 3) / 4;   // words that can hold KS bytes at byte offset <= 3
  static constexpr int kGB = KS == 3 ? 7 : 3;   // 42 = 6 x 7 groups, 18 = 6 x 3 groups
  static_assert(GPS % kGB == 0, "batch size must divide the groups of a stage");
  template <bool kFull, int j0, int i>
  static __device__ __forceinline__ void load(uint32_t (&g)[kGB][NW + 1], uint32_t& rp, uint32_t pitch, uint32_t cs, int& kh,
                                              int groups_left) {
    if constexpr (i < kGB) {
      constexpr int j = j0 + i;
      const bool on = kFull || j < groups_left;      // kFull: every group of the stage lies inside the tensor
#pragma unroll
      for (int q = 0; q < NW; ++q) g[i][q] = on ? lds32(rp + 4 * q) : 0u;
      g[i][NW] = 0u;
      if constexpr (GPS % KS == 0) {          // stage-aligned channels (3x3): compile-time pattern
        rp += ((j % KS) == KS - 1) ? (cs - (KS - 1) * pitch) : pitch;
      } else {
        ++kh;
        if (kh == KS) { kh = 0; rp += cs - (KS - 1) * pitch; } else { rp += pitch; }
      }
      load<kFull, j0, i + 1>(g, rp, pitch, cs, kh, groups_left);
    }
  }
  template <int j0, int i>
  static __device__ __forceinline__ void merge(uint32_t (&w)[36], const uint32_t (&g)[kGB][NW + 1], uint32_t sh8, uint32_t xcol) {
    if constexpr (i < kGB) {
      constexpr int j = j0 + i;
      conv_extract<KS, j, 0, NW + 1>(w, g[i], sh8);
      stream_flush<(j * KS) / 14, ((j + 1) * KS) / 14>(w, xcol);
      merge<j0, i + 1>(w, g, sh8, xcol);
    }
  }
  template <bool kFull, int j0>
  static __device__ __forceinline__ void run_(uint32_t (&w)[36], uint32_t xcol, uint32_t& rp, uint32_t sh8, uint32_t pitch,
                                              uint32_t cs, int& kh, int groups_left) {
    if constexpr (j0 < GPS) {
      uint32_t g[kGB][NW + 1];
      load<kFull, j0, 0>(g, rp, pitch, cs, kh, groups_left);
      merge<j0, 0>(w, g, sh8, xcol);
      run_<kFull, j0 + kGB>(w, xcol, rp, sh8, pitch, cs, kh, groups_left);
    }
  }
  template <int j0>
  static __device__ __forceinline__ void run(uint32_t (&w)[36], uint32_t xcol, uint32_t& rp, uint32_t sh8, uint32_t pitch,
                                             uint32_t cs, int& kh, int groups_left) {
    if (groups_left >= GPS) run_<true, j0>(w, xcol, rp, sh8, pitch, cs, kh, groups_left);
    else run_<false, j0>(w, xcol, rp, sh8, pitch, cs, kh, groups_left);
  }
};

// ---- GEMM: this thread's 144-byte window (9 x 16 B, the row's K range rounded down to 16) -> 9 tiles.
template <int SHIFT>
__device__ __forceinline__ void gemm_restride(const uint32_t (&r)[37], uint32_t xcol) {
#pragma unroll
  for (int t = 0; t < kChunkTiles; ++t) {
    uint32_t o[4];
#pragma unroll
    for (int jw = 0; jw < 4; ++jw) {
      const int byte0 = SHIFT + 14 * t + 4 * jw;             // compile-time after unrolling
      const int wi = byte0 >> 2, bs = byte0 & 3;
      uint32_t v = bs == 0 ? r[wi] : __byte_perm(r[wi], r[wi + 1], 0x3210u + 0x1111u * bs);
      if (jw == 3) v &= 0xffffu;                             // bytes 14, 15 of the tile are padding
      o[jw] = v;
    }
    tmem_st4(xcol + t * 4, o[0], o[1], o[2], o[3]);
  }
}

// ---- fused epilogue of one block-row (14 channels) for this thread's activation row.
enum EpiKind { kEpiI8 = 0, kEpiI8ResFast = 1, kEpiI8Res = 2, kEpiI32 = 3, kEpiGeneric = 4 };
struct EpiCtx {
  int flags;
  bool row_ok;
  int64_t out_base;     // element offset of (this row, channel 0)
  int64_t cs;           // channel stride in elements
  int relu_lo;          // 0 when ReLU acts on the accumulator, else INT_MIN
  int out_lo;           // 0 when ReLU acts on the int8 result, else -128
  uint32_t sat;
  int lane;
};

template <int KIND, bool SAT>
__device__ __forceinline__ void epilogue_row(const TcParams& p, EpiCtx& ec, uint32_t (&v)[16], int cb, int n_ok,
                                             const float* sc, const int32_t* bi);

// One activation stage by TMA: GEMM = rows [m0, m0+128) x 144 bytes from the 16-byte aligned column at or
// before the stage's first k; conv = one [planes][rows][pitch] box per image the tile touches.
template <int MODE>
__device__ __forceinline__ void tma_stage(const TcLaunch& L, uint32_t dst, uint64_t* bar, int chunk, int64_t m0,
                                          uint32_t n_seg, uint32_t img0, int ih_first) {
  const TcParams& p = L.p;
  if constexpr (MODE == kModeGemm) {
    mbar_arrive_expect_tx(bar, static_cast<uint32_t>(kGemmSlotBytes));
    tma_load_2d(dst, &L.tmap, (chunk * kChunkTiles * kBlock) & ~15, static_cast<int>(m0), bar);
  } else {
    constexpr int KS = MODE == kModeConv7 ? 7 : 3;
    mbar_arrive_expect_tx(bar, n_seg * static_cast<uint32_t>(p.box_bytes));
    const int c_first = (chunk * (126 / KS)) / KS;
    for (uint32_t sg = 0; sg < n_seg; ++sg)
      tma_load_4d(dst + sg * p.seg_bytes, &L.tmap, -p.halo_lpad, sg == 0 ? ih_first : -p.pad, c_first,
                  static_cast<int>(img0 + sg), bar);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) bsr_tc_kernel(const __grid_constant__ TcLaunch L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcParams& p = L.p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* x_full = bars;                      // [2]  count 4 (one elected lane per producer warp of the half)
  uint64_t* x_empty = bars + 2;                 // [2]  count kIssuers
  uint64_t* w_full = bars + 4;                  // [kWStages] count 1 (+tx bytes)
  uint64_t* w_empty = w_full + kWStages;        // [kWStages] count kIssuers
  uint64_t* acc_full = w_empty + kWStages;      // [1]  count kIssuers
  uint64_t* h_full = acc_full + 1;              // [kMaxRingSlots] count 32 (loader lanes)
  uint64_t* h_empty = h_full + kMaxRingSlots;   // [kMaxRingSlots] count 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_empty + kMaxRingSlots);
  float* s_scale = reinterpret_cast<float*>(smem + kSmemScale);
  int32_t* s_bias = reinterpret_cast<int32_t*>(smem + kSmemBias);
  int64_t* s_rowoff = reinterpret_cast<int64_t*>(smem + kSmemRowOff);
  int32_t* s_hbase = reinterpret_cast<int32_t*>(smem + kSmemHBase);
  int32_t* s_outrow = reinterpret_cast<int32_t*>(smem + kSmemOutRow);

  constexpr bool kRing = (MODE != kModeDirect);
  constexpr bool kHalo = (MODE == kModeConv3 || MODE == kModeConv7);
  constexpr int KS = MODE == kModeConv7 ? 7 : 3;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t gi = blockIdx.x % L.n_groups;
  const int64_t mtile = blockIdx.x / L.n_groups;
  const int64_t m0 = mtile * kTileM;
  const uint32_t g_br0 = L.groups[gi].br0_rows & 0xffffu, g_rows = L.groups[gi].br0_rows >> 16;
  const uint32_t g_bb = L.groups[gi].batch_begin, g_be = L.groups[gi].batch_end;
  long long* tl = p.timeline ? p.timeline + static_cast<size_t>(blockIdx.x) * 32 : nullptr;
  if (tl && threadIdx.x == 0) tl[0] = clock64();

  // ---- prologue, two block-wide barriers: (A) barriers / TMEM allocation / tables, (B) accumulators zeroed
  int n_out_rows = 0;
  uint32_t R0 = 0;
  if constexpr (kHalo) {
    // which input rows does this tile need?  Output rows R0.. (global row id n*Ho + oh); consecutive output rows
    // of one image share input rows, so halo row ids advance by `stride`; a new image starts a fresh set.
    R0 = static_cast<uint32_t>(m0) / static_cast<uint32_t>(p.Wo);
    const uint32_t m_last = static_cast<uint32_t>(m0 + kTileM - 1 < p.M ? m0 + kTileM - 1 : p.M - 1);
    n_out_rows = static_cast<int>(m_last / static_cast<uint32_t>(p.Wo) - R0) + 1;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], 4); mbar_init(&x_empty[s], kIssuers); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kIssuers); }
    mbar_init(acc_full, kIssuers);
    // h_full: 32 loader lanes arrive (cp.async / register path), or one arrival + transaction bytes (TMA)
    if (!p.use_tma)
      for (int s = 0; s < kMaxRingSlots; ++s) { mbar_init(&h_full[s], 32); mbar_init(&h_empty[s], 4); }
    fence_mbar_init();
  } else if (kHalo && !p.use_tma && threadIdx.x == 32) {
    int hb = 0;
    uint32_t n = R0 / static_cast<uint32_t>(p.Ho);
    uint32_t oh = R0 - n * p.Ho;
    for (int rr = 0; rr < n_out_rows; ++rr) {
      s_hbase[rr] = hb;
      s_outrow[rr] = static_cast<int32_t>(n);
      if (++oh == static_cast<uint32_t>(p.Ho)) { oh = 0; ++n; hb += max(p.ksz, p.stride); } else { hb += p.stride; }
    }
    s_hbase[kMaxOutRows - 1] = s_hbase[n_out_rows - 1] + p.ksz;   // rows in use
  }
  if (warp == kProducerWarps) {
    tmem_alloc_dyn(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  // TMA loader: owns the ring barriers and starts filling the ring before the rest of the CTA is set up
  uint32_t n_seg = 1, img0 = 0;
  int ih_first = 0;
  if (kRing && p.use_tma && warp == kWarpALoad) {
    if constexpr (kHalo) {
      img0 = R0 / static_cast<uint32_t>(p.Ho);
      const uint32_t m_last = static_cast<uint32_t>(m0 + kTileM - 1 < p.M ? m0 + kTileM - 1 : p.M - 1);
      n_seg = m_last / static_cast<uint32_t>(p.Ho * p.Wo) - img0 + 1;
      ih_first = static_cast<int>(R0 - img0 * p.Ho) * p.stride - p.pad;
    }
    if (elect_one()) {
      for (int s = 0; s < kMaxRingSlots; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], 4); }
      fence_mbar_init();
      uint32_t step = 0;
      for (uint32_t b = g_bb; b < g_be && step < static_cast<uint32_t>(p.ring_slots); ++b) {
        const uint32_t bw = L.batches[b];
        if (!(bw & kBatchFirst)) continue;
        tma_stage<MODE>(L, smem_u32(smem + kSmemRing) + step * p.slot_bytes, &h_full[step], static_cast<int>(bw >> 16), m0,
                        n_seg, img0, ih_first);
        ++step;
      }
    }
    __syncwarp();
  }
  // per-channel epilogue constants of this group
  if (threadIdx.x < g_rows * kBlock) {
    const int c = g_br0 * kBlock + threadIdx.x;
    const bool ok = c < p.epi.n_channels;
    s_scale[threadIdx.x] = (ok && p.epi.chan_scale) ? p.epi.chan_scale[c] : 0.f;
    s_bias[threadIdx.x] = (ok && p.epi.bias) ? p.epi.bias[c] : 0;
  }
  if (kHalo && !p.use_tma) {
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem + kSmemRing);
    const int ring_words = (p.ring_slots * p.slot_bytes + kRingSlack) >> 2;
    for (int i = threadIdx.x; i < ring_words; i += kThreads) ring[i] = 0u;   // pads must read as zero
    for (int i = threadIdx.x; i < kMaxHaloRows; i += kThreads) s_rowoff[i] = -1;
  }
  tc_fence_before();
  __syncthreads();                                                             // (A)
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  int hr_used = 0;
  if (kHalo && !p.use_tma) {
    hr_used = s_hbase[kMaxOutRows - 1];
    for (int i = threadIdx.x; i < n_out_rows * p.ksz; i += kThreads) {
      const int rr = i / p.ksz, kh = i - rr * p.ksz;
      const int64_t n = s_outrow[rr];
      const int oh = static_cast<int>(static_cast<int64_t>(R0 + rr) - n * p.Ho);
      const int ih = oh * p.stride - p.pad + kh;
      if (static_cast<unsigned>(ih) < static_cast<unsigned>(p.H))
        s_rowoff[s_hbase[rr] + kh] = ((n * p.C) * p.H + ih) * static_cast<int64_t>(p.Wp);
    }
  }
  if (warp < 4) {  // zero the accumulators: every MMA accumulates, so issue order across warps is free
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    for (uint32_t c = 0; c < g_rows * kTile; c += 4) tmem_st4(tmem_base + lane_base + c, 0u, 0u, 0u, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();                                                             // (B)
  tc_fence_after();

  const uint32_t ring_addr = smem_u32(smem + kSmemRing);
  if (tl && threadIdx.x == 0) tl[1] = clock64();

  if (warp < kProducerWarps) {
    // =================================================================== activation producers
    const int half = warp >> 2;                 // owns stages with (step & 1) == half
    const int tid = threadIdx.x & 127;          // activation row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t xcol = tmem_base + lane_base + kAccCols + half * kXStageCols;
    const int64_t m = m0 + tid;
    const bool row_ok = m < p.M;
    // conv: this thread's output position
    int ow = 0, oh = 0;
    uint32_t img = 0;
    uint32_t thr_off = 0, sh8 = 0;
    if (MODE != kModeGemm && (kHalo || p.conv)) {
      const uint32_t mm = static_cast<uint32_t>(row_ok ? m : m0);
      const uint32_t R = mm / static_cast<uint32_t>(p.Wo);
      ow = static_cast<int>(mm - R * p.Wo);
      img = R / static_cast<uint32_t>(p.Ho);
      oh = static_cast<int>(R - img * p.Ho);
      if constexpr (kHalo) {
        const uint32_t xb = static_cast<uint32_t>(ow * p.stride - p.pad + p.halo_lpad);
        if (p.use_tma) {
          // one box per image: rows start at the first input row of the image's first output row in this tile
          const uint32_t img0 = R0 / static_cast<uint32_t>(p.Ho);
          const uint32_t seg = img - img0;
          const int ohf = seg == 0 ? static_cast<int>(R0 - img0 * p.Ho) : 0;
          thr_off = seg * static_cast<uint32_t>(p.seg_bytes) + static_cast<uint32_t>((oh - ohf) * p.stride) * p.halo_pitch + (xb & ~3u);
        } else {
          const int rr = static_cast<int>(R - R0);
          thr_off = static_cast<uint32_t>(s_hbase[rr]) * p.halo_pitch + (xb & ~3u);
        }
        sh8 = (xb & 3u) * 8u;
      }
    }
    uint32_t step = 0;
    for (uint32_t b = g_bb; b < g_be; ++b) {
      const uint32_t bw = L.batches[b];
      if (!(bw & kBatchFirst)) continue;        // one activation stage per K chunk
      const bool my = (step & 1u) == static_cast<uint32_t>(half);
      const uint32_t use = step >> 1;
      const uint32_t slot = step % static_cast<uint32_t>(kRing ? p.ring_slots : 1);
      const uint32_t sphase = (step / static_cast<uint32_t>(kRing ? p.ring_slots : 1)) & 1u;
      ++step;
      if (!my) continue;
      const int chunk = static_cast<int>(bw >> 16);
      const int k_chunk0 = chunk * kChunkTiles * kBlock;
      if (tl && threadIdx.x == 0 && step == 3) tl[8] = clock64();
      if constexpr (kRing) mbar_wait(&h_full[slot], sphase);
      if (tl && threadIdx.x == 0 && step == 3) tl[9] = clock64();
      mbar_wait(&x_empty[half], (use & 1u) ^ 1u);
      tc_fence_after();
      if (tl && threadIdx.x == 0 && step == 1) tl[2] = clock64();
      if (tl && threadIdx.x == 0 && step == 3) tl[10] = clock64();
      if constexpr (MODE == kModeGemm) {
        // ---------------- GEMM rows from the ring: 9 x 16 B (pitch 144 B: conflict-free 16-byte reads)
        const uint32_t src = ring_addr + slot * p.slot_bytes + tid * 144;
        uint32_t r[37];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const uint4 v = lds128(src + i * 16);
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        }
        r[36] = 0u;
        switch (k_chunk0 & 15) {                 // 126 * chunk mod 16: even values only
          case 0: gemm_restride<0>(r, xcol); break;
          case 2: gemm_restride<2>(r, xcol); break;
          case 4: gemm_restride<4>(r, xcol); break;
          case 6: gemm_restride<6>(r, xcol); break;
          case 8: gemm_restride<8>(r, xcol); break;
          case 10: gemm_restride<10>(r, xcol); break;
          case 12: gemm_restride<12>(r, xcol); break;
          default: gemm_restride<14>(r, xcol); break;
        }
      } else if constexpr (kHalo) {
        // ---------------- conv from the staged input rows
        constexpr int GPS = 126 / KS;
        const uint32_t cs = static_cast<uint32_t>(p.halo_rows) * p.halo_pitch;
        const int g0 = chunk * GPS;                                   // first (c, kh) group of the stage
        const int c_first = g0 / KS;
        int kh = g0 - c_first * KS;
        const int groups_left = p.C * KS - g0;                        // groups past the tensor contribute zeros
        uint32_t rp = ring_addr + slot * p.slot_bytes + thr_off + kh * p.halo_pitch;
        uint32_t w[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) w[i] = 0u;
        ConvBatch<KS>::template run<0>(w, xcol, rp, sh8, p.halo_pitch, cs, kh, groups_left);
      } else if (!p.conv) {
        // ---------------- DIRECT GEMM rows (unaligned base / leading dimension)
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
      } else {
        // ---------------- DIRECT conv: any k x k, incremental (c, kh, kw) walk; a tile's 14 loads are independent
        int k = k_chunk0;
        const int kk = p.ksz * p.ksz;
        int c = k / kk;
        int rem = k - c * kk;
        int kh = rem / p.ksz, kw = rem - kh * p.ksz;
        const int8_t* im = p.x + static_cast<int64_t>(img) * p.C * p.H * p.Wp;
        const int ih0 = oh * p.stride - p.pad, iw0 = ow * p.stride - p.pad;
#pragma unroll 1
        for (int t = 0; t < kChunkTiles; ++t) {
          uint32_t v[kBlock];
#pragma unroll
          for (int i = 0; i < kBlock; ++i) {
            const int ih = ih0 + kh, iw = iw0 + kw;
            v[i] = 0u;
            if (row_ok && k < p.K && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H) &&
                static_cast<unsigned>(iw) < static_cast<unsigned>(p.W))
              v[i] = static_cast<uint8_t>(im[(static_cast<int64_t>(c) * p.H + ih) * p.Wp + iw]);
            ++k;
            if (++kw == p.ksz) { kw = 0; if (++kh == p.ksz) { kh = 0; ++c; } }
          }
          uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int i = 0; i < kBlock; ++i) w[i >> 2] |= v[i] << ((i & 3) * 8);
          tmem_st4(xcol + t * 4, w[0], w[1], w[2], w[3]);
        }
      }
      if (tl && threadIdx.x == 0 && step == 3) tl[11] = clock64();
      tmem_st_wait();
      if (tl && threadIdx.x == 0 && step == 3) tl[12] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kRing) mbar_arrive(&h_empty[slot]);
        mbar_arrive(&x_full[half]);
      }
      if (tl && threadIdx.x == 0 && step == 1) tl[3] = clock64();
      if (tl && threadIdx.x == 0 && step == 3) tl[13] = clock64();
    }

    // =================================================================== epilogue (same 8 warps)
    if (tl && threadIdx.x == 0) tl[4] = clock64();
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tl && threadIdx.x == 0) tl[5] = clock64();
    EpiCtx ec;
    ec.flags = p.epi.flags;
    ec.row_ok = row_ok;
    ec.out_base = 0;
    if (row_ok) {
      const int64_t im = m / p.lay.rows_per_image;
      const int64_t pix = m - im * p.lay.rows_per_image;
      if (p.lay.row_len > 0) {
        const int64_t r = pix / p.lay.row_len;
        ec.out_base = im * p.lay.image_stride + r * p.lay.row_pitch + (pix - r * p.lay.row_len) * p.lay.row_stride;
      } else {
        ec.out_base = im * p.lay.image_stride + pix * p.lay.row_stride;
      }
    }
    ec.cs = p.lay.chan_stride;
    ec.relu_lo = (ec.flags & ACCEL_RELU) ? 0 : INT_MIN;
    ec.out_lo = (ec.flags & ACCEL_RELU_OUT) ? 0 : -128;
    ec.sat = 0;
    ec.lane = lane;
    // one uniform decision per CTA: which specialised path handles every block-row of this tile
    const bool sat_on = p.epi.sat_count != nullptr;
    int kind = kEpiGeneric;
    if (!p.epi.chan_absmax && (ec.flags & ACCEL_OUT_I8))
      kind = !p.epi.residual ? kEpiI8 : (p.res_fast ? kEpiI8ResFast : kEpiI8Res);
    else if (!p.epi.chan_absmax && !sat_on && (ec.flags & ACCEL_OUT_I32)) kind = kEpiI32;
    if (sat_on && kind != kEpiGeneric) kind += 8;
    for (uint32_t g = half; g < g_rows; g += 2) {
      uint32_t v[16];
      if (tl && threadIdx.x == 0 && g < 6) tl[14 + 3 * (g >> 1)] = clock64();
      tmem_ld16(tmem_base + lane_base + g * kTile, v);
      const int cb = (g_br0 + g) * kBlock;
      const int n_ok = min(kBlock, p.epi.n_channels - cb);   // channels of this block-row that exist (warp-uniform)
      const float* sc = s_scale + g * kBlock;
      const int32_t* bi = s_bias + g * kBlock;
      switch (kind) {
        case kEpiI8: epilogue_row<kEpiI8, false>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI8ResFast: epilogue_row<kEpiI8ResFast, false>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI8Res: epilogue_row<kEpiI8Res, false>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI8 + 8: epilogue_row<kEpiI8, true>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI8ResFast + 8: epilogue_row<kEpiI8ResFast, true>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI8Res + 8: epilogue_row<kEpiI8Res, true>(p, ec, v, cb, n_ok, sc, bi); break;
        case kEpiI32: epilogue_row<kEpiI32, false>(p, ec, v, cb, n_ok, sc, bi); break;
        default: epilogue_row<kEpiGeneric, false>(p, ec, v, cb, n_ok, sc, bi); break;
      }
      if (tl && threadIdx.x == 0 && g < 6) tl[15 + 3 * (g >> 1)] = clock64();
    }
    if (tl && threadIdx.x == 0) tl[23] = clock64();
    if (p.epi.sat_count) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, ec.sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
    tc_fence_before();
    if (tl && threadIdx.x == 0) tl[6] = clock64();
  } else if (warp < kWarpWLoad) {
    // =================================================================== MMA issuers (uniform datapath)
    const uint32_t me = static_cast<uint32_t>(warp - kProducerWarps);
    if (elect_one()) {
      const uint32_t idesc0 = idesc_i8(kTileM, 0);
      const uint32_t w_addr = smem_u32(smem + kSmemW);
      const uint64_t bdesc0 = smem_desc_kmajor(0, 128, 256);
      const uint32_t bdesc_hi = static_cast<uint32_t>(bdesc0 >> 32);
      const uint32_t bdesc_lo0 = static_cast<uint32_t>(bdesc0);
      uint32_t step = 0, wcount = 0, xs = 0;
      uint32_t opb = L.groups[gi].op_begin;
      for (uint32_t b = g_bb; b < g_be; ++b) {
        const uint32_t bw = L.batches[b];
        if (bw & kBatchFirst) {
          xs = step & 1u;
          mbar_wait(&x_full[xs], (step >> 1) & 1u);
          ++step;
        }
        const uint32_t ws_i = wcount % kWStages;
        mbar_wait(&w_full[ws_i], (wcount / kWStages) & 1u);
        tc_fence_after();
        const uint32_t n_runs = bw & 0xffu;
        const uint32_t xa = tmem_base + kAccCols + xs * kXStageCols;
        const uint32_t blo = bdesc_lo0 | (((w_addr + ws_i * kWStageBytes) >> 4) & 0x3FFFu);
        for (uint32_t i = me; i < n_runs; i += kIssuers) {
          const OpRec o = L.ops[opb + i];
          const uint32_t d = tmem_base + (o.d_n & 0x1ffu);
          const uint32_t idesc = idesc0 | (o.d_n & 0x7E0000u);
          const uint32_t a = xa + (o.a_b & 0x1ffu);
          const uint64_t bdesc = (static_cast<uint64_t>(bdesc_hi) << 32) | (blo + (o.a_b >> 16));
          mma_i8_ts(d, a, bdesc, idesc, 1u);
        }
        opb += n_runs;
        mma_commit(&w_empty[ws_i]);
        if (bw & kBatchLast) mma_commit(&x_empty[xs]);
        if (b == g_be - 1) mma_commit(acc_full);
        ++wcount;
      }
      if (g_bb == g_be) mbar_arrive(acc_full);  // nothing stored in this group: outputs are bias only
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == kWarpWLoad) {
    // =================================================================== weight loader
    if (lane == 0) {
      uint32_t wcount = 0;
      const uint8_t* src = p.blob + static_cast<size_t>(L.groups[gi].blob_off16) * 16;
      for (uint32_t b = g_bb; b < g_be; ++b) {
        const uint32_t bw = L.batches[b];
        const uint32_t ws_i = wcount % kWStages;
        mbar_wait(&w_empty[ws_i], ((wcount / kWStages) & 1u) ^ 1u);
        const uint32_t bytes = (((bw >> 8) & 0x3fu) + 1u) * kBTileBytes;
        mbar_arrive_expect_tx(&w_full[ws_i], bytes);
        bulk_g2s(smem + kSmemW + ws_i * kWStageBytes, src, bytes, &w_full[ws_i]);
        src += bytes;
        ++wcount;
      }
    }
    __syncwarp();
  } else if (kRing) {
    // =================================================================== activation loader
    if (p.use_tma) {
      // One elected thread, one TMA tensor tile per stage (GEMM) or per image the tile touches (conv); rows,
      // columns, channels and images outside the tensor arrive as zeros - that is the convolution's padding.
      if (elect_one()) {
        uint32_t step = 0;
        for (uint32_t b = g_bb; b < g_be; ++b) {
          const uint32_t bw = L.batches[b];
          if (!(bw & kBatchFirst)) continue;
          const uint32_t slot = step % static_cast<uint32_t>(p.ring_slots);
          const uint32_t sphase = (step / static_cast<uint32_t>(p.ring_slots)) & 1u;
          ++step;
          if (step <= static_cast<uint32_t>(p.ring_slots)) continue;      // issued in the prologue
          mbar_wait(&h_empty[slot], sphase ^ 1u);
          tma_stage<MODE>(L, ring_addr + slot * p.slot_bytes, &h_full[slot], static_cast<int>(bw >> 16), m0, n_seg, img0,
                          ih_first);
        }
      }
      __syncwarp();
    } else {
    // A single warp issues ~1 dependent instruction per 5 cycles, so everything that does not change from stage
    // to stage is computed once per tile and kept in registers: a copy then costs an add, a compare and the LDGSTS.
    constexpr int kLd = kHalo ? 8 : 9;
    uint32_t l_src[kLd];      // conv: byte offset of the piece inside channel plane 0 of the stage; GEMM: low word of the row offset
    uint32_t l_meta[kLd];     // conv: smem offset / 4 | bytes << 16, 0xffffffff = nothing to copy
    [[maybe_unused]] uint32_t l_src_hi[kHalo ? 1 : 9];
    bool l_fast = false;
    if constexpr (kHalo) {
      // a lane owns up to 8 (input row, 16- or 4-byte piece) pairs and copies them for every channel plane
      const int items = hr_used * p.halo_ipr;
      const int piece = p.halo_vec == 16 ? 16 : 4;
      l_fast = p.halo_vec != 0 && items <= 32 * kLd && p.x_small;
#pragma unroll
      for (int i = 0; i < kLd; ++i) {
        const int idx = lane + 32 * i;
        l_src[i] = 0; l_meta[i] = 0xffffffffu;
        if (l_fast && idx < items) {
          const int h = p.halo_ipr > 1 ? static_cast<int>(__umulhi(static_cast<unsigned>(idx), p.ipr_magic)) : idx;
          const int wd = idx - h * p.halo_ipr;
          const int64_t off = s_rowoff[h];
          if (off >= 0) {      // rows outside the image are never written: the ring was zeroed in the prologue
            const int nb = min(piece, p.W - wd * piece);
            const uint32_t d = static_cast<uint32_t>(h * p.halo_pitch + p.halo_lpad + wd * piece);
            l_src[i] = static_cast<uint32_t>(off + wd * piece);
            l_meta[i] = (d >> 2) | (static_cast<uint32_t>(nb) << 16);
          }
        }
      }
    } else {
      // GEMM: item = it * 32 + lane covers (row = item / 9, 16-byte piece = item % 9); items 288 apart are 32 rows apart
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int item = i * 32 + lane;
        const int row = item / 9, ch = item - row * 9;
        const int64_t o = (m0 + row) * p.lda + ch * 16;
        l_src[i] = static_cast<uint32_t>(o);
        l_src_hi[i] = static_cast<uint32_t>(o >> 32);
        l_meta[i] = static_cast<uint32_t>(row * 144 + ch * 16) | (static_cast<uint32_t>(row) << 16);
      }
    }
    uint32_t step = 0;
    for (uint32_t b = g_bb; b < g_be; ++b) {
      const uint32_t bw = L.batches[b];
      if (!(bw & kBatchFirst)) continue;
      const uint32_t slot = step % static_cast<uint32_t>(p.ring_slots);
      const uint32_t sphase = (step / static_cast<uint32_t>(p.ring_slots)) & 1u;
      ++step;
      if (tl && lane == 0 && step <= 2) tl[24 + 2 * (step - 1)] = clock64();
      mbar_wait(&h_empty[slot], sphase ^ 1u);
      const int chunk = static_cast<int>(bw >> 16);
      uint8_t* dst_slot = smem + kSmemRing + slot * p.slot_bytes;
      if constexpr (MODE == kModeGemm) {
        const int a0 = (chunk * kChunkTiles * kBlock) & ~15;
        const uint32_t slot_addr = ring_addr + slot * p.slot_bytes;
        const int64_t row32 = 32 * p.lda;
        if (a0 + 144 <= p.K && m0 + kTileM <= p.M) {          // interior: no clamping
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int8_t* g = p.x + ((static_cast<int64_t>(l_src_hi[i]) << 32) | l_src[i]) + a0;
            const uint32_t d = slot_addr + (l_meta[i] & 0xffffu);
#pragma unroll
            for (int r = 0; r < 4; ++r) cp_async16_zfill_s(d + r * (32 * 144), g + r * row32, 16);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const int ch16 = static_cast<int>((l_meta[i] & 0xffffu) % 144u);
            int nbytes = p.K - (a0 + ch16);
            nbytes = nbytes > 16 ? 16 : (nbytes < 0 ? 0 : nbytes);
            const int8_t* g = p.x + ((static_cast<int64_t>(l_src_hi[i]) << 32) | l_src[i]) + a0;
            const uint32_t d = slot_addr + (l_meta[i] & 0xffffu);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const bool ok = m0 + static_cast<int64_t>(l_meta[i] >> 16) + 32 * r < p.M && nbytes > 0;
              cp_async16_zfill_s(d + r * (32 * 144), ok ? g + r * row32 : p.x, ok ? nbytes : 0);
            }
          }
        }
        cp_async_mbar_arrive(&h_full[slot]);
      } else {
        constexpr int GPS = 126 / KS;
        const int g0 = chunk * GPS;
        const int c_first = g0 / KS;
        const int ipr = p.halo_ipr;                                   // copy items per input row
        const int items = hr_used * ipr;
        const uint32_t cs = static_cast<uint32_t>(p.halo_rows) * p.halo_pitch;
        const int64_t HW = static_cast<int64_t>(p.H) * p.Wp;
        const int nch = min(p.halo_nch, p.C - c_first);
        if (l_fast) {
          const int8_t* base = p.x + c_first * HW;
          const uint32_t slot_addr = ring_addr + slot * p.slot_bytes;
#pragma unroll
          for (int i = 0; i < kLd; ++i) {
            if (l_meta[i] == 0xffffffffu) continue;
            const int8_t* g = base + l_src[i];
            const uint32_t d = slot_addr + ((l_meta[i] & 0xffffu) << 2);
            const int nb = static_cast<int>(l_meta[i] >> 16);
            if (p.halo_vec == 16) {
#pragma unroll
              for (int c = 0; c < 14; ++c)
                if (c < nch) cp_async16_zfill_s(d + c * cs, g + c * HW, nb);
            } else {
#pragma unroll
              for (int c = 0; c < 14; ++c)
                if (c < nch) cp_async4_zfill_s(d + c * cs, g + c * HW, 4);
            }
          }
        } else
        for (int idx = lane; idx < items; idx += 32) {
          const int h = ipr > 1 ? static_cast<int>(__umulhi(static_cast<unsigned>(idx), p.ipr_magic)) : idx;
          const int wd = idx - h * ipr;
          const int64_t off = s_rowoff[h];
          if (p.halo_vec == 16) {
            // 16-byte aligned rows: 16-byte copies, the row tail is zero-filled by the copy itself
            uint8_t* d = dst_slot + h * p.halo_pitch + 16 + wd * 16;
            const int nbytes = off >= 0 ? min(16, p.W - wd * 16) : 0;
            const int8_t* g = off >= 0 ? p.x + off + c_first * HW + wd * 16 : p.x;
            for (int c = 0; c < p.halo_nch; ++c) {
              cp_async16_zfill(d, (nbytes && c < nch) ? g : p.x, c < nch ? nbytes : 0);
              g += HW;
              d += cs;
            }
          } else if (p.halo_vec == 4) {
            uint8_t* d = dst_slot + h * p.halo_pitch + 4 + wd * 4;
            const int8_t* g = off >= 0 ? p.x + off + c_first * HW + wd * 4 : p.x;
            for (int c = 0; c < p.halo_nch; ++c) {
              cp_async4_zfill(d, (off >= 0 && c < nch) ? g : p.x, off >= 0 && c < nch);
              g += HW;
              d += cs;
            }
          } else {
            // rows at any byte alignment: aligned 32-bit loads, funnel shift, mask the row tail.  Channel planes
            // are handled in batches so that the independent loads of a batch are all in flight together.
            uint8_t* d = dst_slot + h * p.halo_pitch + 4 + wd * 4;
            const int nb = min(4, p.W - wd * 4);
            const uint32_t mask = nb >= 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u);
            constexpr int kB = 7;
            for (int c0 = 0; c0 < p.halo_nch; c0 += kB) {
              uint32_t lo[kB], hi[kB], sh[kB];
#pragma unroll
              for (int i = 0; i < kB; ++i) {
                const int c = c0 + i;
                lo[i] = 0u; hi[i] = 0u; sh[i] = 0u;
                if (off >= 0 && c < nch) {
                  const uintptr_t a = reinterpret_cast<uintptr_t>(p.x + off + (c_first + c) * HW + wd * 4);
                  const uint32_t* a4 = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
                  const uint32_t s = static_cast<uint32_t>(a & 3);
                  sh[i] = s * 8;
                  lo[i] = __ldg(a4);
                  if (s + nb > 4) hi[i] = __ldg(a4 + 1);
                }
              }
#pragma unroll
              for (int i = 0; i < kB; ++i) {
                if (c0 + i < p.halo_nch)
                  *reinterpret_cast<uint32_t*>(d + (c0 + i) * cs) = __funnelshift_r(lo[i], hi[i], sh[i]) & mask;
              }
            }
          }
        }
        if (tl && lane == 0 && step <= 2) tl[28 + (step - 1)] = clock64();
        if (p.halo_vec) cp_async_mbar_arrive(&h_full[slot]);
        else mbar_arrive(&h_full[slot]);
        if (tl && lane == 0 && step <= 2) tl[25 + 2 * (step - 1)] = clock64();
      }
    }
    cp_async_wait_all();
    }
  }

  __syncthreads();
  if (tl && threadIdx.x == 0) tl[7] = clock64();
  if (warp == kProducerWarps) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, kTmemCols);
  }
}

template <int KIND, bool SAT>
__device__ __forceinline__ void epilogue_row(const TcParams& p, EpiCtx& ec, uint32_t (&v)[16], int cb, int n_ok,
                                             const float* sc, const int32_t* bi) {
  const int64_t o0 = ec.out_base + static_cast<int64_t>(cb) * ec.cs;
  if constexpr (KIND == kEpiI8 || KIND == kEpiI8ResFast || KIND == kEpiI8Res) {
    // golden_models.cpp:378-411 per channel: one float32 multiply, round-half-even, saturate;
    // then (residual variants) add_residual_int8, golden_models.cpp:465-490: (main*s_main + res*s_res) / s_out
    constexpr bool kRes = KIND != kEpiI8;
    int8_t* o = reinterpret_cast<int8_t*>(p.out) + o0;
    int rv[kBlock];
    if constexpr (kRes) {
      const int8_t* __restrict__ r = p.epi.residual + o0;
#pragma unroll
      for (int h = 0; h < kBlock; ++h) rv[h] = (ec.row_ok && h < n_ok) ? static_cast<int>(r[h * ec.cs]) : 0;
    }
    float scl[kBlock];
    int bia[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) { scl[h] = sc[h]; bia[h] = bi[h]; }
    tmem_ld_wait();
    int q[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) {
      const int acc = max(static_cast<int>(v[h]) + bia[h], ec.relu_lo);
      const float f = __fmul_rn(__int2float_rn(acc), scl[h]);
      int r8 = cvt_sat_s8(f);
      if constexpr (SAT) {
        // round-half-even leaves [-128, 127] exactly when f >= 127.5 (-> 128) or f < -128.5 (-128.5 -> -128 stays)
        ec.sat += (ec.row_ok && h < n_ok && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
      }
      if constexpr (kRes) {
        const float a = __fmul_rn(__int2float_rn(r8), p.epi.res_scale_main);
        const float r = __fmul_rn(__int2float_rn(rv[h]), p.epi.res_scale_res);
        const float s = __fadd_rn(a, r);
        float d;
        if constexpr (KIND == kEpiI8ResFast) {   // correctly rounded for every (int8, int8) pair: verified on the host
          const float q0 = __fmul_rn(s, p.res_rcp);
          const float e = __fmaf_rn(-q0, p.epi.res_scale_out, s);
          d = __fmaf_rn(e, p.res_rcp, q0);
        } else {
          d = __fdiv_rn(s, p.epi.res_scale_out);
        }
        r8 = cvt_sat_s8(d);
      }
      q[h] = max(r8, ec.out_lo);                 // relu_int8 (golden_models.cpp:278-283) when requested
    }
    if (!ec.row_ok || (p.dbg_flags & 1)) return;
    if (n_ok == kBlock) {
#pragma unroll
      for (int h = 0; h < kBlock; ++h) o[h * ec.cs] = static_cast<int8_t>(q[h]);
    } else {
#pragma unroll
      for (int h = 0; h < kBlock; ++h)
        if (h < n_ok) o[h * ec.cs] = static_cast<int8_t>(q[h]);
    }
  } else if constexpr (KIND == kEpiI32) {
    int bia[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) bia[h] = bi[h];
    tmem_ld_wait();
    int acc[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) acc[h] = max(static_cast<int>(v[h]) + bia[h], ec.relu_lo);
    if (!ec.row_ok) return;
    int32_t* o = reinterpret_cast<int32_t*>(p.out) + o0;
    if (n_ok == kBlock && ec.cs == 1 && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
#pragma unroll
      for (int h = 0; h < kBlock; h += 2) *reinterpret_cast<int2*>(o + h) = make_int2(acc[h], acc[h + 1]);
    } else {
#pragma unroll
      for (int h = 0; h < kBlock; ++h)
        if (h < n_ok) o[h * ec.cs] = acc[h];
    }
  } else {
    // every option: saturation counter, per-channel |acc| maximum, float32 output
    const int flags = ec.flags;
    const int8_t* __restrict__ resid = p.epi.residual;
    int rv[kBlock];
    if (resid && ec.row_ok) {
#pragma unroll
      for (int h = 0; h < kBlock; ++h) rv[h] = h < n_ok ? static_cast<int>(resid[o0 + h * ec.cs]) : 0;
    }
    tmem_ld_wait();
    int acc[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) acc[h] = max(static_cast<int>(v[h]) + bi[h], ec.relu_lo);
    if (p.epi.chan_absmax) {
#pragma unroll
      for (int h = 0; h < kBlock; ++h) {
        const int a = ec.row_ok ? (acc[h] < 0 ? (acc[h] == INT_MIN ? INT_MAX : -acc[h]) : acc[h]) : 0;
        const int wmax = __reduce_max_sync(0xffffffffu, a);
        if (ec.lane == 0 && wmax > 0 && h < n_ok) atomicMax(p.epi.chan_absmax + cb + h, wmax);
      }
    }
    if (!ec.row_ok) return;
    if (flags & ACCEL_OUT_I32) {
      int32_t* o = reinterpret_cast<int32_t*>(p.out) + o0;
#pragma unroll
      for (int h = 0; h < kBlock; ++h)
        if (h < n_ok) o[h * ec.cs] = acc[h];
    } else if (flags & ACCEL_OUT_F32) {
      float* o = reinterpret_cast<float*>(p.out) + o0;
#pragma unroll
      for (int h = 0; h < kBlock; ++h)
        if (h < n_ok) o[h * ec.cs] = __fmul_rn(__int2float_rn(acc[h]), sc[h]);
    } else {
      int8_t* o = reinterpret_cast<int8_t*>(p.out) + o0;
#pragma unroll
      for (int h = 0; h < kBlock; ++h) {
        const float f = __fmul_rn(__int2float_rn(acc[h]), sc[h]);
        const int qi = __float2int_rn(f);
        int q = min(127, max(-128, qi));
        ec.sat += (qi != q && h < n_ok) ? 1u : 0u;
        if (resid) {
          const float a = __fmul_rn(__int2float_rn(q), p.epi.res_scale_main);
          const float r = __fmul_rn(__int2float_rn(rv[h]), p.epi.res_scale_res);
          q = cvt_sat_s8(__fdiv_rn(__fadd_rn(a, r), p.epi.res_scale_out));
        }
        q = max(q, ec.out_lo);
        if (h < n_ok) o[h * ec.cs] = static_cast<int8_t>(q);
      }
    }
  }
}

}  // namespace accel
