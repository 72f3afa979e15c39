// Weight-stationary stem: 7x7 / stride 2 / pad 3 convolution of a <= 4-channel image, ReLU / per-channel requant and the
// 3x3 / stride 2 / pad 1 max-pool that follows it, in one kernel (conv2d_int8_im2col + requantize_int32_to_int8 +
// maxpool2d_int8 of golden_models.cpp:883-933, :378-411, :534-571; the ResNet stem of resnet_inference.cpp:61-127).
//
// Same operand roles as conv_ws.cuh: D[co][px] += W[co][k] * X[k][px], lane = output channel, column = conv pixel.
//   * One tile = one conv output row.  Stride 2 in x is removed by splitting every input row into its even and odd
//     columns while it is staged (E[j] = X[2j], O[j] = X[2j+1]): conv column x needs E[x + s] for kw = 3 + 2s
//     (s = -1, 0, 1) and O[x + s] for kw = 4 + 2s (s = -2 .. 1) - a stride-1 problem with seven column shifts.
//   * K = (c, kh): 21 of 32 rows of a chunk; two chunks per tile (E and O), [32 k-rows][128 pixels] with the 128-byte
//     swizzle = one MN-major atom row per k.  Pixels 112..127 are zero (never written).
//   * Column shifts through the accumulator address, which must be an even column: every chunk is staged twice, as it is
//     and shifted right by one pixel (E1[j] = E[j-1], O1[j] = O[j-1]), so that the seven taps all land in ONE accumulator
//     at offsets 0 / 2 / 4: seven MMAs (M = 64, N = 128, K = 32) per tile and one TMEM read per output.
//   * <= 64 output channels: two IMAGES share an accumulator set (lanes 0-15 / 16-31 of every lane quadrant).
//   * Input rows are requested two conv rows ahead of their conversion (16-byte LDGSTS into a raw ring, zero fill above /
//     below the image = the vertical padding); the three converter warps de-interleave them from shared memory (LDS -> byte
//     permutes -> STS into the swizzled operand stage) and never wait for global memory.  (Round 1 loaded with LDG inside
//     the converter - one exposed L2 / HBM latency per stage and warp; a 4-D TMA box [W][7][C] of 224-byte rows was tried
//     and is slower still: ~1 us per 4.7 KB box.)
//   * The pool is taken on the INT32 accumulators: requant (and ReLU) are monotone in the accumulator for a positive
//     per-channel factor, so max-then-requant equals requant-then-max bit for bit, and only one value in four is
//     converted.  An item = a STRIP of G pooled rows of an image pair = conv rows 2*yp0 - 1 .. 2*yp1 - 1; the epilogue
//     keeps the running column maxima of the open pooled row in registers, and an odd conv row 2yp + 1 both closes pooled
//     row yp and opens yp + 1 (its values replace the maxima).  Only the strip's lead-in row is computed twice:
//     (2G + 1) / G conv rows per pooled row instead of 3 (G is chosen on the host for the least work per SM).
//     Saturation is counted per conv output (every row of the strip but the lead-in) against per-channel thresholds.
#pragma once
#include "conv_ws.cuh"

namespace accel {

constexpr int kStEpiWarps = 8;
constexpr int kStWarpIssue = kStEpiWarps;            // 8: MMA issuer
constexpr int kStWarpConv = kStWarpIssue + 1;        // 9..11: converters (LDGSTS -> raw ring -> byte de-interleave -> operand stage)
constexpr int kStConvWarps = 3;
constexpr int kStConvThreads = kStConvWarps * 32;
constexpr int kStThreads = (kStWarpConv + kStConvWarps) * 32;   // 384: 168 registers per thread
constexpr int kStStageBytes = 16384;                 // E, O, E1, O1 chunks of 4 KB
constexpr int kStSlots = 8;
constexpr int kStAhead = 2;                          // conv rows a converter warp has requested beyond the one it converts
constexpr int kStRawSlots = kStConvWarps * kStAhead * 2;      // raw stages (one image's (c, kh) rows of a conv row) in shared memory
constexpr int kStWBytes = 7 * kWsTapBytes;           // one 64(128) x 32 tile per kw
constexpr int kStSmemBar = 1024;
constexpr int kStN = 128;
constexpr uint32_t kStY0 = 2;                        // accumulator column of conv pixel 0
constexpr uint32_t kStAccSets = 3;                   // accumulator sets in TMEM: the issuer runs two conv rows ahead of the epilogue
constexpr uint32_t kStAccCols = 168;                 // columns per set (132 used: 128 pixels + the shifts)
constexpr int kStSmemFixed = 1024 + kStSmemBar + kStWBytes + kStSlots * kStStageBytes + 64;      // + kStRawSlots * raw_slot_bytes; 64 zero bytes

struct StemParams {
  int32_t C, H, W, B;            // input
  int32_t Hc, Wc, Hp, Wp;        // conv output, pooled output (Hc = 2 * Hp)
  int32_t c_out, in_pitch, out_pitch;
  int32_t n_pairs;               // ceil(B / 2)
  int32_t G, n_strips;           // pooled rows per item, items per image pair
  int32_t n_items;               // n_pairs * n_strips
  FastDiv d_strips;
  const uint8_t* wblob;          // [kw][4096]
  accel_epilogue epi;
  int8_t* out;                   // pooled [B][c_out][Hp][out_pitch]
  int32_t chan_stride;           // Hp * out_pitch
  int64_t image_stride;          // c_out * chan_stride
  const int8_t* x;
  int32_t raw_slot_bytes;        // C * 7 * (16 + W + 32) rounded up to 128: one raw stage (rows with zero padding left and right)
  int32_t dbg;                   // developer aid: bit 0 = epilogue does no work, bit 1 = no MMAs
  long long* timeline;           // developer aid: clock64 stamps of CTA 0, 8 kinds x 64 conv rows (tools/stem_timeline.py)
};

// scatter of the stored blocks: reference K index = (c * 7 + kh) * 7 + kw  ->  tile kw, row co, column c * 7 + kh
__global__ void stem_scatter_kernel(const int8_t* __restrict__ blocks, const int32_t* __restrict__ blk_row,
                                    const int32_t* __restrict__ col_idx, int64_t nnz, int32_t c_in, int32_t c_out,
                                    uint8_t* __restrict__ blob) {
  const int64_t b = blockIdx.x;
  if (b >= nnz) return;
  const int br = blk_row[b], bc = col_idx[b];
  for (int i = threadIdx.x; i < kBlock * kBlock; i += blockDim.x) {
    const int h = i / kBlock, w = i - h * kBlock;
    const int co = br * kBlock + h, k = bc * kBlock + w;
    if (co >= c_out || k >= c_in * 49) continue;
    const int kw = k % 7, ckh = k / 7;
    blob[static_cast<size_t>(kw) * kWsTapBytes + (co >> 3) * 256 + (ckh >> 4) * 128 + (co & 7) * 16 + (ckh & 15)] =
        static_cast<uint8_t>(blocks[b * 196 + i]);
  }
}

// conv rows of item `it`: first and last (inclusive), and the strip's first pooled row
__device__ __forceinline__ void stem_item(const StemParams& p, uint32_t it, uint32_t& pr, int& yp0, int& r0, int& r1) {
  pr = fdiv(it, p.d_strips);
  const int strip = static_cast<int>(it - pr * static_cast<uint32_t>(p.n_strips));
  yp0 = strip * p.G;
  const int yp1 = min(yp0 + p.G, p.Hp);
  r0 = max(2 * yp0 - 1, 0);
  r1 = 2 * yp1 - 1;
}

// range of a 16-column chunk against the per-channel clipping thresholds; exact recount only when it trips
__device__ __forceinline__ void stem_count16(const uint32_t (&z)[16], bool count, bool need_min, int hi_a, int lo_a, int col, int Wc,
                                             uint32_t& sat) {
  int amax = INT_MIN, amin = INT_MAX;
#pragma unroll
  for (int e = 0; e < 16; e += 2) amax = __vimax3_s32(amax, static_cast<int>(z[e]), static_cast<int>(z[e + 1]));
  if (need_min) {
#pragma unroll
    for (int e = 0; e < 16; e += 2) amin = __vimin3_s32(amin, static_cast<int>(z[e]), static_cast<int>(z[e + 1]));
  }
  if (count && (amax > hi_a || amin < lo_a)) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int a = static_cast<int>(z[e]);
      sat += (col + e < Wc && (a > hi_a || a < lo_a)) ? 1u : 0u;
    }
  }
}

// Per-thread constants of the stem epilogue, and the running column maxima of one open pooled row: 64 conv columns in four
// chunks of 16 (the unit of a TMEM load) plus the column left of them (the upper half's neighbour).
struct StemEpi {
  int bias, relu_lo, hi_a, lo_a, col0, xp_lo, xp_hi, Wc;
  float sf;
  bool need_min, half;
};
struct StemRow {
  uint32_t c[4][16];
  int left;
};

// even conv row 2yp: fold into the maxima of the open pooled row (TMEM load of chunk k + 1 in flight while chunk k is folded)
__device__ __forceinline__ void stem_row_fold(const StemEpi& e, uint32_t acc, StemRow& st, bool count, uint32_t& sat) {
  uint32_t za[16], zb[16];
  int zl = INT_MIN;
  if (e.half) zl = static_cast<int>(tmem_ld1(acc - 1));
  tmem_ld16(acc, za);
  tmem_ld_wait();
  tmem_ld16(acc + 16, zb);
#define STEM_FOLD(Z, C)                                                                                   \
  stem_count16(Z, count, e.need_min, e.hi_a, e.lo_a, e.col0 + 16 * (C), e.Wc, sat);                       \
  _Pragma("unroll") for (int i = 0; i < 16; ++i)                                                          \
    st.c[C][i] = static_cast<uint32_t>(max(static_cast<int>(st.c[C][i]), static_cast<int>(Z[i])));
  STEM_FOLD(za, 0)
  tmem_ld_wait();
  tmem_ld16(acc + 32, za);
  STEM_FOLD(zb, 1)
  tmem_ld_wait();
  tmem_ld16(acc + 48, zb);
  STEM_FOLD(za, 2)
  tmem_ld_wait();
  STEM_FOLD(zb, 3)
#undef STEM_FOLD
  st.left = max(st.left, zl);
}

// lead-in row of a strip (odd row 2 * yp0 - 1): its values open pooled row yp0
__device__ __forceinline__ void stem_row_open(const StemEpi& e, uint32_t acc, StemRow& nw) {
  nw.left = INT_MIN;
  if (e.half) nw.left = static_cast<int>(tmem_ld1(acc - 1));
  tmem_ld16(acc, nw.c[0]);
  tmem_ld16(acc + 16, nw.c[1]);
  tmem_ld16(acc + 32, nw.c[2]);
  tmem_ld16(acc + 48, nw.c[3]);
  tmem_ld_wait();
}

// odd conv row 2yp + 1: loaded straight into the registers of the NEXT open row (nw); closes pooled row yp held in `old`
// (horizontal 3-max, one requant per pooled value, 4-byte stores) - no register is copied
__device__ __forceinline__ void stem_row_close(const StemEpi& e, uint32_t acc, StemRow& old, StemRow& nw, bool count, bool emit,
                                               int8_t* orow, uint32_t& sat) {
  nw.left = INT_MIN;
  if (e.half) nw.left = static_cast<int>(tmem_ld1(acc - 1));
  tmem_ld16(acc, nw.c[0]);
  tmem_ld_wait();
  int carry = max(old.left, nw.left);          // column col0 - 1 is padding in the lower half: stays INT_MIN
#define STEM_CLOSE(C)                                                                                     \
  {                                                                                                       \
    stem_count16(nw.c[C], count, e.need_min, e.hi_a, e.lo_a, e.col0 + 16 * (C), e.Wc, sat);               \
    _Pragma("unroll") for (int i = 0; i < 16; ++i)                                                        \
      old.c[C][i] = static_cast<uint32_t>(max(static_cast<int>(old.c[C][i]), static_cast<int>(nw.c[C][i]))); \
    if (emit && e.xp_lo + 8 * (C) < e.xp_hi) {                                                            \
      uint32_t qv[8];                                                                                     \
      _Pragma("unroll") for (int k = 0; k < 8; ++k) {                                                     \
        const int m = __vimax3_s32(k == 0 ? carry : static_cast<int>(old.c[C][2 * k - 1]), static_cast<int>(old.c[C][2 * k]), \
                                   static_cast<int>(old.c[C][2 * k + 1]));                                \
        const int a = max(m + e.bias, e.relu_lo);                                                         \
        qv[k] = e.xp_lo + 8 * (C) + k < e.xp_hi ? cvt_sat_s8_raw(__fmul_rn(__int2float_rn(a), e.sf)) : 0u; \
      }                                                                                                   \
      *reinterpret_cast<uint32_t*>(orow + 8 * (C)) = pack4_b0(qv[0], qv[1], qv[2], qv[3]);                \
      if (e.xp_lo + 8 * (C) + 4 < e.xp_hi) *reinterpret_cast<uint32_t*>(orow + 8 * (C) + 4) = pack4_b0(qv[4], qv[5], qv[6], qv[7]); \
    }                                                                                                     \
    carry = static_cast<int>(old.c[C][15]);                                                               \
  }
  tmem_ld16(acc + 16, nw.c[1]);
  STEM_CLOSE(0)
  tmem_ld_wait();
  tmem_ld16(acc + 32, nw.c[2]);
  STEM_CLOSE(1)
  tmem_ld_wait();
  tmem_ld16(acc + 48, nw.c[3]);
  STEM_CLOSE(2)
  tmem_ld_wait();
  STEM_CLOSE(3)
#undef STEM_CLOSE
}

__global__ void __launch_bounds__(kStThreads, 1) stem_ws_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_full = bars;                       // [kStSlots] count 1 (lane 0 of the converter warp that owns the row)
  uint64_t* a_empty = a_full + kStSlots;         // [kStSlots] count 1 (tcgen05.commit)
  uint64_t* w_full = a_empty + kStSlots;         // [1]
  uint64_t* acc_full = w_full + 1;               // [kStAccSets] count 1
  uint64_t* acc_empty = acc_full + kStAccSets;   // [kStAccSets] count kStEpiWarps
  uint64_t* raw_full = acc_empty + kStAccSets;            // [kStRawSlots] count 32 (cp.async.mbarrier.arrive of the owning warp's lanes)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_full + kStRawSlots);
  const uint32_t w_addr = base + kStSmemBar;
  const uint32_t a_addr = w_addr + kStWBytes;    // 1024-aligned: 28672 = 28 * 1024
  uint8_t* raw = smem + kStSmemBar + kStWBytes + kStSlots * kStStageBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t n_items = static_cast<uint32_t>(p.n_items);

  long long* const tl = (ACCEL_DEV && p.timeline && blockIdx.x == 0) ? p.timeline : nullptr;      // -DACCEL_DEV=1 builds only
#define STEM_STAMP(KIND, ROW) if (tl && (ROW) >= 32u && (ROW) < 96u) tl[(KIND) * 64 + ((ROW) - 32u)] = clock64();
  griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStSlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < kStAccSets; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kStEpiWarps); }
    for (int s = 0; s < kStRawSlots; ++s) mbar_init(&raw_full[s], 32);
    fence_mbar_init();
  }
  // the operand ring starts out zero: pixels 112..127 and k-rows 21..31 of every chunk are never written again
  for (uint32_t i = threadIdx.x; i < kStSlots * kStStageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + kStSmemBar + kStWBytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = threadIdx.x; i < (static_cast<uint32_t>(kStRawSlots) * p.raw_slot_bytes + 64u) / 16u; i += blockDim.x)
    reinterpret_cast<uint4*>(raw)[i] = make_uint4(0u, 0u, 0u, 0u);      // raw ring: the row padding stays zero for good
  fence_proxy_async_smem();
  if (warp == kStWarpIssue) {
    tmem_alloc_dyn(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kStEpiWarps) {
    // =================================================================== epilogue: thread = (image of the pair, channel)
    griddep_wait();
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int sub = lane >> 4;
    const int co = q * 16 + (lane & 15);
    const bool ch_ok = co < p.c_out;
    const float sf = ch_ok ? p.epi.chan_scale[co] : 0.f;
    const int bias = (ch_ok && p.epi.bias) ? p.epi.bias[co] : 0;
    const bool relu = (p.epi.flags & ACCEL_RELU) != 0;
    const int relu_lo = relu ? 0 : INT_MIN;
    const bool sat_on = p.epi.sat_count != nullptr;
    // clipped  <=>  max(a + bias, relu_lo) outside [lo_c, hi_c]  <=>  a > hi_a or a < lo_a  (a = raw accumulator)
    int lo_a = INT_MIN, hi_a = INT_MAX;
    if (sat_on && ch_ok) {
      int lo_c, hi_c;
      ws_sat_bounds(sf, lo_c, hi_c);
      const long long h = static_cast<long long>(hi_c) - bias, l = static_cast<long long>(lo_c) - bias;
      hi_a = hi_c < relu_lo ? INT_MIN : static_cast<int>(min(max(h, static_cast<long long>(INT_MIN)), static_cast<long long>(INT_MAX)));
      lo_a = relu_lo >= lo_c ? INT_MIN : static_cast<int>(min(max(l, static_cast<long long>(INT_MIN)), static_cast<long long>(INT_MAX)));
    }
    StemEpi e;
    e.bias = bias; e.relu_lo = relu_lo; e.hi_a = hi_a; e.lo_a = lo_a; e.sf = sf; e.Wc = p.Wc;
    e.need_min = !relu;                             // with ReLU nothing clips on the low side (lo_a stays INT_MIN)
    e.half = half != 0;
    e.col0 = half ? 64 : 0;                         // first conv column this thread loads (64 columns) and counts
    e.xp_lo = half ? 32 : 0; e.xp_hi = half ? p.Wp : min(32, p.Wp);       // pooled columns this thread produces
    uint32_t sat = 0, nrow = 0;                     // nrow: conv rows seen by this CTA (accumulator set = nrow & 1)
    // one conv row: wait for its accumulator set, run BODY (acc = TMEM address of this thread's first column), hand the set back
#define STEM_ROW(BODY)                                                                                    \
  {                                                                                                       \
    const uint32_t ab = nrow % kStAccSets;                                                                \
    const uint32_t acc = tmem_base + lane_base + ab * kStAccCols + kStY0 + static_cast<uint32_t>(e.col0); \
    mbar_wait(&acc_full[ab], (nrow / kStAccSets) & 1u);                                                   \
    tc_fence_after();                                                                                     \
    if (threadIdx.x == 0) STEM_STAMP(4, nrow)                                                             \
    ++nrow;                                                                                               \
    BODY;                                                                                                 \
    tc_fence_before();                                                                                    \
    __syncwarp();                                                                                         \
    if (threadIdx.x == 0) STEM_STAMP(5, nrow - 1u)                                                        \
    if (lane == 0) mbar_arrive(&acc_empty[ab]);                                                           \
  }
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      uint32_t pr;
      int yp0, r0, r1;
      stem_item(p, it, pr, yp0, r0, r1);
      if (ACCEL_DEV && (p.dbg & 1)) {                              // developer aid: hand the accumulators straight back
        for (int yc = r0; yc <= r1; ++yc) STEM_ROW((void)acc)
        continue;
      }
      const int yp1 = (r1 + 1) >> 1;
      const int img = static_cast<int>(2u * pr) + sub;
      const bool ok = img < p.B && ch_ok;           // threads without an image / a channel compute along and store nothing
      const bool count = sat_on && ok;
      int8_t* orow = p.out + static_cast<int64_t>(img < p.B ? img : 0) * p.image_stride + static_cast<int64_t>(ch_ok ? co : 0) * p.chan_stride +
                     static_cast<int64_t>(yp0) * p.out_pitch + e.xp_lo;
      // Two register sets take turns as "open pooled row": an odd conv row is loaded into the idle set while it closes the
      // other one, so the maxima never move between registers (as one loop-carried array they cost ~200 moves per row).
      // Only B is carried around the loop.
      StemRow B;
      B.left = INT_MIN;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 16; ++i) B.c[c][i] = 0x80000000u;
      if (r0 & 1) STEM_ROW(stem_row_open(e, acc, B))                       // lead-in: counted by the strip above
      int yp = yp0;
      for (; yp + 1 < yp1; yp += 2) {
        StemRow A;
        STEM_ROW(stem_row_fold(e, acc, B, count, sat))
        STEM_ROW(stem_row_close(e, acc, B, A, count, ok, orow, sat))
        STEM_ROW(stem_row_fold(e, acc, A, count, sat))
        STEM_ROW(stem_row_close(e, acc, A, B, count, ok, orow + p.out_pitch, sat))
        orow += 2 * p.out_pitch;
      }
      if (yp < yp1) {
        StemRow A;
        STEM_ROW(stem_row_fold(e, acc, B, count, sat))
        STEM_ROW(stem_row_close(e, acc, B, A, count, ok, orow, sat))
      }
    }
#undef STEM_ROW
    if (sat_on) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
  } else if (warp == kStWarpIssue) {
    // =================================================================== weights (once) + MMA issue
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, kStWBytes);
      for (int i = 0; i < 7; ++i) bulk_g2s(smem + kStSmemBar + i * kWsTapBytes, p.wblob + i * kWsTapBytes, kWsTapBytes, w_full);
      mbar_wait(w_full, 0u);
      const uint32_t idesc = idesc_i8_bmn(64u, static_cast<uint32_t>(kStN));
      const uint64_t adesc0 = smem_desc_kmajor(0, 128, 256);
      const uint64_t bdesc0 = smem_desc_any(0, 1024, 1024, 2u);           // MN-major, 128-byte swizzle, k-groups 1 KB apart
      const uint32_t a_hi = static_cast<uint32_t>(adesc0 >> 32), b_hi = static_cast<uint32_t>(bdesc0 >> 32);
      const uint32_t a_lo0 = static_cast<uint32_t>(adesc0) | ((w_addr >> 4) & 0x3FFFu);
      const uint32_t b_lo0 = static_cast<uint32_t>(bdesc0);
      uint32_t as = 0, aph = 0, nrow = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        uint32_t pr;
        int yp0, r0, r1;
        stem_item(p, it, pr, yp0, r0, r1);
        const uint32_t n_sub = (2u * pr + 1u < static_cast<uint32_t>(p.B)) ? 2u : 1u;
        for (int yc = r0; yc <= r1; ++yc) {
          const uint32_t ab = nrow % kStAccSets;
          mbar_wait(&acc_empty[ab], ((nrow / kStAccSets) & 1u) ^ 1u);
          tc_fence_after();
          STEM_STAMP(6, nrow)
          ++nrow;
          for (uint32_t sub = 0; sub < n_sub; ++sub) {
            const uint32_t z = tmem_base + ab * kStAccCols + ((sub * 16u) << 16);
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            if (sub == 0) STEM_STAMP(2, nrow - 1u)
            const uint32_t e_lo = b_lo0 | (((a_addr + as * kStStageBytes) >> 4) & 0x3FFFu);
            const uint64_t be = (static_cast<uint64_t>(b_hi) << 32) | e_lo, bo = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 256u);
            const uint64_t be1 = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 512u), bo1 = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 768u);
            auto wt = [&](int kw) { return (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + kw * (kWsTapBytes >> 4)); };
            if (!(ACCEL_DEV && (p.dbg & 2))) {
              // conv pixel x lives in column 2 + x.  Unshifted chunks: shift 0 -> +2, shift -2 -> +4.  Chunks shifted right by
              // one pixel: shift -1 -> +2, shift +1 -> +0.  The first MMA overwrites [2, 130); columns 0, 1, 130, 131 are never read.
              mma_i8_ss(z + 2, wt(3), be, idesc, 0u);      // E,  kw 3, shift  0
              mma_i8_ss(z + 2, wt(1), be1, idesc, 1u);     // E1, kw 1, shift -1
              mma_i8_ss(z + 0, wt(5), be1, idesc, 1u);     // E1, kw 5, shift +1
              mma_i8_ss(z + 2, wt(4), bo, idesc, 1u);      // O,  kw 4, shift  0
              mma_i8_ss(z + 4, wt(0), bo, idesc, 1u);      // O,  kw 0, shift -2
              mma_i8_ss(z + 2, wt(2), bo1, idesc, 1u);     // O1, kw 2, shift -1
              mma_i8_ss(z + 0, wt(6), bo1, idesc, 1u);     // O1, kw 6, shift +1
            }
            mma_commit(&a_empty[as]);
            if (++as == kStSlots) { as = 0; aph ^= 1u; }
          }
          mma_commit(&acc_full[ab]);
          STEM_STAMP(3, nrow - 1u)
        }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // =================================================================== converters: raw rows -> E / O / E1 / O1 chunks
    // A conv row (the stages of both images of the pair) belongs to converter warp (row number % 3): a lane owns up to five
    // 32-byte units of the (c, kh) rows (32 input bytes -> 16 even + 16 odd pixels).  The raw bytes of the warp's NEXT
    // kStAhead rows are already on their way (16-byte LDGSTS into the warp's own raw ring; rows above / below the image are
    // zero-filled: that is the vertical padding) while it converts the current one, so no global-memory latency is exposed.
    // fence.proxy.async is paid once per row and warp, and the three warps overlap each other's.
    const int lw = warp - kStWarpConv;
    const int upr = p.W >> 5;                        // 32-byte units per input row (W % 32 == 0, <= 7)
    // Raw stage: (c, kh) rows of pitch 16 + W + 32 - sixteen zero bytes on the left, 32 on the right (zeroed once, the copies
    // only write the W real bytes).  Conversion walks upr + 1 units per row: the last one reads the right padding, which
    // makes the "one more pixel" of the shifted copies fall out of the same code; units beyond the last row read a zero
    // area and store zeros into k-rows 21.. of the chunk, which are zero anyway.  No predicates in the per-stage code.
    const int rp = 16 + p.W + 32;
    const int n_load = p.C * 7 * upr, n_conv = p.C * 7 * (upr + 1);
    constexpr int kLoadUnits = 5;                    // ceil(160 / 32): C * 7 * (W / 32) <= 160
    constexpr int kConvUnits = 6;                    // ceil(C * 7 * (upr + 1) / 32) <= ceil(28 * 8 / 32) + ...: checked on the host
    uint32_t soff[kConvUnits], roff[kConvUnits];     // roff: offset in the raw stage, or 0x80000000 | zero-area offset
    uint32_t ldst[kLoadUnits];
    int32_t goff[kLoadUnits], ukh[kLoadUnits];       // ukh: kh of the unit's row, or a value no row test passes
#pragma unroll
    for (int k = 0; k < kLoadUnits; ++k) {
      const int un = lane + 32 * k;
      const bool has = un < n_load;
      const int rowid = has ? un / upr : 0, u = has ? un - rowid * upr : 0;
      const int c = rowid / 7, kh = rowid - c * 7;
      ldst[k] = static_cast<uint32_t>(rowid * rp + 16 + 32 * u);
      goff[k] = (c * p.H + kh - 3) * p.in_pitch + 32 * u;       // + 2 * yc * pitch
      ukh[k] = has ? kh : (1 << 20);
    }
    const uint32_t zero_addr = smem_u32(raw) + static_cast<uint32_t>(kStRawSlots) * static_cast<uint32_t>(p.raw_slot_bytes);      // 64 zero bytes
#pragma unroll
    for (int k = 0; k < kConvUnits; ++k) {
      const int un = lane + 32 * k;
      const int rowid = un / (upr + 1), u = un - rowid * (upr + 1);
      // units beyond the last row: zeros into k-row 31, which no (c, kh) uses (C * 7 <= 28)
      soff[k] = un < n_conv ? static_cast<uint32_t>(rowid) * 128u + ((static_cast<uint32_t>(u) ^ (rowid & 7)) << 4)
                            : 31u * 128u + (static_cast<uint32_t>(lane & 7) << 4);
      roff[k] = un < n_conv ? static_cast<uint32_t>(rowid * rp + 16 + 32 * u) : 0x80000000u;
    }
    const int n_conv_iter = (n_conv + 31) >> 5;
    const uint32_t raw_slot = static_cast<uint32_t>(p.raw_slot_bytes);
    const uint32_t raw_w = smem_u32(raw) + static_cast<uint32_t>(lw) * (2u * kStAhead) * raw_slot;      // this warp's ring
    uint64_t* raw_full_w = raw_full + lw * (2 * kStAhead);
    // cursor over the conv rows of this CTA, kStAhead owned rows ahead of the one being converted
    uint32_t c_it = blockIdx.x, c_pr = 0, c_nsub = 1, c_ridx = 0, c_own = 0;      // c_own: owned rows requested so far
    int c_yc = 0, c_r1 = -1;
    bool c_valid = c_it < n_items;
    auto cursor_item = [&]() {
      int yp0, r0;
      stem_item(p, c_it, c_pr, yp0, r0, c_r1);
      c_yc = r0;
      c_nsub = (2u * c_pr + 1u < static_cast<uint32_t>(p.B)) ? 2u : 1u;
    };
    auto cursor_next_row = [&]() {
      ++c_ridx;
      if (++c_yc > c_r1) { c_it += gridDim.x; c_valid = c_it < n_items; if (c_valid) cursor_item(); }
    };
    auto request_row = [&]() {                       // LDGSTS of the cursor's row (must be an owned one), then on to the next owned row
      for (uint32_t sub = 0; sub < c_nsub; ++sub) {
        const uint32_t slot = (c_own % kStAhead) * 2u + sub;
        const int8_t* src = p.x + static_cast<int64_t>(2u * c_pr + sub) * p.C * p.H * p.in_pitch + static_cast<int64_t>(2 * c_yc) * p.in_pitch;
        const uint32_t dst = raw_w + slot * raw_slot;
#pragma unroll
        for (int k = 0; k < kLoadUnits; ++k) {
          const bool in = static_cast<uint32_t>(2 * c_yc + ukh[k] - 3) < static_cast<uint32_t>(p.H);      // false for lanes without a unit
          const int8_t* g = src + (in ? goff[k] : 0);
          if (ukh[k] < 8) {
            cp_async16_zfill_s(dst + ldst[k], g, in ? 16 : 0);
            cp_async16_zfill_s(dst + ldst[k] + 16u, g + 16, in ? 16 : 0);
          }
        }
        cp_async_mbar_arrive(&raw_full_w[slot]);
      }
      ++c_own;
      cursor_next_row();
      while (c_valid && c_ridx % kStConvWarps != static_cast<uint32_t>(lw)) cursor_next_row();
    };
    griddep_wait();
    if (c_valid) {
      cursor_item();
      while (c_valid && c_ridx % kStConvWarps != static_cast<uint32_t>(lw)) cursor_next_row();
      for (int i = 0; i < kStAhead && c_valid; ++i) request_row();
    }
    uint32_t sidx = 0, ridx = 0, own = 0;            // running stage / conv row numbers (identical in the issuer), owned rows converted
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      uint32_t pr;
      int yp0, r0, r1;
      stem_item(p, it, pr, yp0, r0, r1);
      const uint32_t n_sub = (2u * pr + 1u < static_cast<uint32_t>(p.B)) ? 2u : 1u;
      for (int yc = r0; yc <= r1; ++yc, sidx += n_sub, ++ridx) {
        if (ridx % kStConvWarps != static_cast<uint32_t>(lw)) continue;
        for (uint32_t sub = 0; sub < n_sub; ++sub) {
          const uint32_t s = sidx + sub, as = s % kStSlots;
          const uint32_t rslot = (own % kStAhead) * 2u + sub;
          mbar_wait(&raw_full_w[rslot], (own / kStAhead) & 1u);
          if (lane == 0 && sub == 0) STEM_STAMP(0, ridx)
          const uint32_t src = raw_w + rslot * raw_slot;
          uint8_t* dst = smem + kStSmemBar + kStWBytes + as * kStStageBytes;
          // all loads of the stage first (up to 18 in flight), then the byte work, then the stores
          uint4 lo[kConvUnits], hi[kConvUnits];
          uint32_t pv[kConvUnits];                     // the four input bytes left of the unit (its last two feed the shifted copies)
#pragma unroll
          for (int k = 0; k < kConvUnits; ++k) {
            if (k >= n_conv_iter) break;               // uniform
            const uint32_t a = (roff[k] & 0x80000000u) ? zero_addr + 16u : src + roff[k];
            lo[k] = lds128(a);
            hi[k] = lds128(a + 16u);
            pv[k] = lds32(a - 4u);
          }
          mbar_wait(&a_empty[as], ((s / kStSlots) & 1u) ^ 1u);
#pragma unroll
          for (int k = 0; k < kConvUnits; ++k) {
            if (k >= n_conv_iter) break;               // uniform
            const uint4 ev = make_uint4(__byte_perm(lo[k].x, lo[k].y, 0x6420), __byte_perm(lo[k].z, lo[k].w, 0x6420),
                                        __byte_perm(hi[k].x, hi[k].y, 0x6420), __byte_perm(hi[k].z, hi[k].w, 0x6420));
            const uint4 od = make_uint4(__byte_perm(lo[k].x, lo[k].y, 0x7531), __byte_perm(lo[k].z, lo[k].w, 0x7531),
                                        __byte_perm(hi[k].x, hi[k].y, 0x7531), __byte_perm(hi[k].z, hi[k].w, 0x7531));
            // shifted right by one pixel: byte 0 comes from the unit on the left (the zero padding at the image edge)
            const uint4 ev1 = make_uint4(__byte_perm(pv[k], ev.x, 0x6542), __funnelshift_l(ev.x, ev.y, 8), __funnelshift_l(ev.y, ev.z, 8),
                                         __funnelshift_l(ev.z, ev.w, 8));
            const uint4 od1 = make_uint4(__byte_perm(pv[k], od.x, 0x6543), __funnelshift_l(od.x, od.y, 8), __funnelshift_l(od.y, od.z, 8),
                                         __funnelshift_l(od.z, od.w, 8));
            *reinterpret_cast<uint4*>(dst + soff[k]) = ev;
            *reinterpret_cast<uint4*>(dst + soff[k] + 4096) = od;
            *reinterpret_cast<uint4*>(dst + soff[k] + 8192) = ev1;
            *reinterpret_cast<uint4*>(dst + soff[k] + 12288) = od1;
          }
        }
        fence_proxy_async_smem();                    // generic-proxy stores -> visible to tcgen05.mma
        __syncwarp();                                // ... and every lane is done reading this row's raw slots
        if (lane == 0)
          for (uint32_t sub = 0; sub < n_sub; ++sub) mbar_arrive(&a_full[(sidx + sub) % kStSlots]);
        if (lane == 0) STEM_STAMP(1, ridx)
        ++own;
        if (c_valid) request_row();                  // into the raw slots just read
        if (lane == 0) STEM_STAMP(7, ridx)
      }
    }
  }

#undef STEM_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == kStWarpIssue) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
}

}  // namespace accel
