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
//   * Column shifts through the accumulator address, which must be an even column: the loaders (a register path anyway)
//     stage every chunk twice, as it is and shifted right by one pixel (E1[j] = E[j-1], O1[j] = O[j-1]), so that the
//     seven taps all land in ONE accumulator at offsets 0 / 2 / 4: seven MMAs (M = 64, N = 128, K = 32) per tile and
//     one TMEM read per output (TMEM reads, 64 B/clk/SM, are what bounds these epilogues: a second accumulator for the
//     odd shifts, as conv_ws.cuh uses, doubles them).
//   * <= 64 output channels: two IMAGES share an accumulator set (lanes 0-15 / 16-31 of every lane quadrant).
//   * The pool is taken on the INT32 accumulators: requant (and ReLU) are monotone in the accumulator for a positive
//     per-channel factor, so max-then-requant equals requant-then-max bit for bit, and only one value in four is
//     converted.  An item = one pooled row of an image pair = conv rows 2yp-1, 2yp, 2yp+1 (rows are recomputed across
//     items: 1.5x the MMAs, which are cheap here); the epilogue keeps the running maximum of its 64 columns in registers.
//     Saturation is still counted per conv output (rows 2yp and 2yp+1 of every item) against per-channel thresholds.
#pragma once
#include "conv_ws.cuh"

namespace accel {

constexpr int kStEpiWarps = 8;
constexpr int kStWarpIssue = kStEpiWarps;            // 8
constexpr int kStWarpLoad = kStWarpIssue + 1;        // 9..14: loaders (LDG -> byte de-interleave -> STS)
constexpr int kStLoadWarps = 3;
constexpr int kStLoadThreads = kStLoadWarps * 32;
constexpr int kStThreads = (kStWarpLoad + kStLoadWarps) * 32;   // 384: 168 registers per thread
constexpr int kStStageBytes = 16384;                 // E, O, E1, O1 chunks of 4 KB
constexpr int kStSlots = 8;
constexpr int kStWBytes = 7 * kWsTapBytes;           // one 64(128) x 32 tile per kw
constexpr int kStSmemBar = 1024;
constexpr int kStN = 128;
constexpr uint32_t kStY0 = 2;                        // accumulator column of conv pixel 0

struct StemParams {
  int32_t C, H, W, B;            // input
  int32_t Hc, Wc, Hp, Wp;        // conv output, pooled output
  int32_t c_out, in_pitch, out_pitch;
  int32_t n_pairs;               // ceil(B / 2)
  int32_t n_items;               // n_pairs * Hp
  FastDiv d_hp;
  const int8_t* x;
  const uint8_t* wblob;          // [kw][4096]
  accel_epilogue epi;
  int8_t* out;                   // pooled [B][c_out][Hp][out_pitch]
  int32_t chan_stride;           // Hp * out_pitch
  int64_t image_stride;          // c_out * chan_stride
  int32_t dbg;                   // developer aid: bit 0 = epilogue does no work, bit 1 = no MMAs, bit 2 = loaders load nothing
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

__global__ void __launch_bounds__(kStThreads, 1) stem_ws_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_full = bars;                       // [kStSlots] count 32 (the loader warp that owns the stage)
  uint64_t* a_empty = a_full + kStSlots;         // [kStSlots] count 1
  uint64_t* w_full = a_empty + kStSlots;         // [1]
  uint64_t* acc_full = w_full + 1;               // [2] count 1
  uint64_t* acc_empty = acc_full + 2;            // [2] count kStEpiWarps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const uint32_t w_addr = base + kStSmemBar;
  const uint32_t a_addr = w_addr + kStWBytes;    // 1024-aligned: 28672 = 28 * 1024

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t n_items = static_cast<uint32_t>(p.n_items);

  griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStSlots; ++s) { mbar_init(&a_full[s], 32); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kStEpiWarps); }
    fence_mbar_init();
  }
  // the ring starts out zero: pixels 112..127 and k-rows 21..31 of every chunk are never written again
  for (uint32_t i = threadIdx.x; i < kStSlots * kStStageBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem + kStSmemBar + kStWBytes)[i] = make_uint4(0u, 0u, 0u, 0u);
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
    const int relu_lo = (p.epi.flags & ACCEL_RELU) ? 0 : INT_MIN;
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
    const int col0 = half ? 64 : 0;                 // first conv column this thread loads (64 columns) and counts
    const int xp_lo = half ? 32 : 0, xp_hi = half ? p.Wp : min(32, p.Wp);       // pooled columns this thread produces
    const bool warp_has_ch = q * 16 < p.c_out;
    uint32_t sat = 0, nrow = 0;                     // nrow: conv rows seen by this CTA (accumulator set = nrow & 1)
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint32_t pr = fdiv(it, p.d_hp);
      const int yp = static_cast<int>(it - pr * static_cast<uint32_t>(p.Hp));
      const int img = static_cast<int>(2u * pr) + sub;
      const bool img_ok = img < p.B;
      int vm[64];                                   // running maximum of conv columns col0 .. col0 + 63
      int vleft = INT_MIN;                          // ... and of column col0 - 1 (the upper half's left neighbour)
#pragma unroll
      for (int j = 0; j < 64; ++j) vm[j] = INT_MIN;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yc = 2 * yp + dy;
        if (yc < 0 || yc >= p.Hc) continue;         // uniform: such rows are not computed at all
        const uint32_t ab = nrow & 1u;
        const uint32_t acc = tmem_base + lane_base + ab * kWsAccCols;
        mbar_wait(&acc_full[ab], (nrow >> 1) & 1u);
        tc_fence_after();
        ++nrow;
        if (warp_has_ch && !(p.dbg & 1)) {
          const bool count = sat_on && dy >= 0 && img_ok && ch_ok;
          // four steps of 16 columns folded into the running maxima.  Clipping is counted the way conv_ws.cuh does it: the
          // chunk's accumulator range against per-channel thresholds, an exact recount only when it trips.
          if (half) {
            const uint32_t zl = tmem_ld1(acc + kStY0 + col0 - 1);
            tmem_ld_wait();
            vleft = max(vleft, static_cast<int>(zl));
          }
#define STEM_FOLD(Z, J0)                                                                            \
  {                                                                                                 \
    int amax = INT_MIN, amin = INT_MAX;                                                             \
    _Pragma("unroll") for (int e = 0; e < 16; ++e) {                                                \
      const int a = static_cast<int>(Z[e]);                                                         \
      vm[(J0) + e] = max(vm[(J0) + e], a);                                                          \
      amax = max(amax, a);                                                                          \
      amin = min(amin, a);                                                                          \
    }                                                                                               \
    if (count && (amax > hi_a || amin < lo_a)) {                                                    \
      _Pragma("unroll") for (int e = 0; e < 16; ++e) {                                              \
        const int a = static_cast<int>(Z[e]);                                                       \
        sat += (col0 + (J0) + e < p.Wc && (a > hi_a || a < lo_a)) ? 1u : 0u;                        \
      }                                                                                             \
    }                                                                                               \
  }
          {
            uint32_t za[16], zb[16];                 // the TMEM load of step k + 1 is in flight while step k is folded
            tmem_ld16(acc + kStY0 + col0, za);
            tmem_ld_wait();
            tmem_ld16(acc + kStY0 + col0 + 16, zb);
            STEM_FOLD(za, 0)
            tmem_ld_wait();
            tmem_ld16(acc + kStY0 + col0 + 32, za);
            STEM_FOLD(zb, 16)
            tmem_ld_wait();
            tmem_ld16(acc + kStY0 + col0 + 48, zb);
            STEM_FOLD(za, 32)
            tmem_ld_wait();
            STEM_FOLD(zb, 48)
          }
#undef STEM_FOLD
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
      }
      // horizontal 3-max (columns outside [0, Wc) are padding), then one requant per pooled value
      if (warp_has_ch && img_ok && ch_ok && !(p.dbg & 1)) {
        int8_t* orow = p.out + static_cast<int64_t>(img) * p.image_stride + static_cast<int64_t>(co) * p.chan_stride +
                       static_cast<int64_t>(yp) * p.out_pitch;
#pragma unroll
        for (int g = 0; g < 8; ++g) {               // 4 pooled outputs per group
          uint32_t qv[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int xp = xp_lo + 4 * g + b;
            constexpr int kDummy = 0;
            const int j = 2 * (4 * g + b) + kDummy;                // vm index of conv column 2 * xp (same in both halves)
            int m = vm[j];                                         // 2 * xp < Wc whenever xp < Wp
            m = max(m, j > 0 ? vm[j > 0 ? j - 1 : 0] : vleft);     // column -1 is padding: vleft stays INT_MIN in the lower half
            if (2 * xp + 1 < p.Wc) m = max(m, vm[j + 1]);
            const int a = max(m + bias, relu_lo);
            qv[b] = xp < xp_hi ? cvt_sat_s8_raw(__fmul_rn(__int2float_rn(a), sf)) : 0u;
          }
          if (xp_lo + 4 * g < xp_hi) *reinterpret_cast<uint32_t*>(orow + xp_lo + 4 * g) = pack4_b0(qv[0], qv[1], qv[2], qv[3]);
        }
      }
    }
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
        const uint32_t pr = fdiv(it, p.d_hp);
        const int yp = static_cast<int>(it - pr * static_cast<uint32_t>(p.Hp));
        const uint32_t n_sub = (2u * pr + 1u < static_cast<uint32_t>(p.B)) ? 2u : 1u;
        for (int dy = -1; dy <= 1; ++dy) {
          const int yc = 2 * yp + dy;
          if (yc < 0 || yc >= p.Hc) continue;
          const uint32_t ab = nrow & 1u;
          mbar_wait(&acc_empty[ab], ((nrow >> 1) & 1u) ^ 1u);
          tc_fence_after();
          ++nrow;
          for (uint32_t sub = 0; sub < n_sub; ++sub) {
            const uint32_t z = tmem_base + ab * kWsAccCols + ((sub * 16u) << 16);
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            const uint32_t e_lo = b_lo0 | (((a_addr + as * kStStageBytes) >> 4) & 0x3FFFu);
            const uint64_t be = (static_cast<uint64_t>(b_hi) << 32) | e_lo, bo = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 256u);
            const uint64_t be1 = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 512u), bo1 = (static_cast<uint64_t>(b_hi) << 32) | (e_lo + 768u);
            auto wt = [&](int kw) { return (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + kw * (kWsTapBytes >> 4)); };
            if (!(p.dbg & 2)) {
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
        }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // =================================================================== loaders: 32 input bytes -> 16 even + 16 odd
    // Stage s (one conv row of one image) belongs to loader warp s % 3: three stages are in flight, and every lane has
    // its (up to) five 32-byte units of a stage outstanding at once.
    griddep_wait();
    const int lw = warp - kStWarpLoad;
    const int upr = p.W >> 5;                        // 32-byte units per input row (W % 32 == 0)
    const int n_units = p.C * 7 * upr;
    constexpr int kUnits = 5;                        // ceil(160 / 32): C * 7 * (W / 32) <= 160
    uint32_t soff[kUnits], stail[kUnits];
    int32_t goff[kUnits], ukh[kUnits], ufirst[kUnits];     // ufirst: 1 = first unit of its row, 2 = last, 0 = inner (3 = both)
#pragma unroll
    for (int k = 0; k < kUnits; ++k) {
      const int un = lane + 32 * k;
      const bool has = un < n_units;
      const int rowid = has ? un / upr : 0, u = has ? un - rowid * upr : 0;
      const int c = rowid / 7, kh = rowid - c * 7;
      soff[k] = static_cast<uint32_t>(rowid) * 128u + ((static_cast<uint32_t>(u) ^ (rowid & 7)) << 4);
      stail[k] = static_cast<uint32_t>(rowid) * 128u + ((static_cast<uint32_t>(u + 1) ^ (rowid & 7)) << 4);
      goff[k] = (c * p.H + kh - 3) * p.in_pitch + 32 * u;       // + 2 * yc * pitch
      ukh[k] = has ? kh : -1000;
      ufirst[k] = (u == 0 ? 1 : 0) | (u == upr - 1 ? 2 : 0);
    }
    uint32_t sidx = 0;                               // running stage number (identical in the issuer)
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint32_t pr = fdiv(it, p.d_hp);
      const int yp = static_cast<int>(it - pr * static_cast<uint32_t>(p.Hp));
      const uint32_t n_sub = (2u * pr + 1u < static_cast<uint32_t>(p.B)) ? 2u : 1u;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yc = 2 * yp + dy;
        if (yc < 0 || yc >= p.Hc) continue;
        for (uint32_t sub = 0; sub < n_sub; ++sub, ++sidx) {
          if (sidx % kStLoadWarps != static_cast<uint32_t>(lw)) continue;
          const uint32_t as = sidx % kStSlots, aph = (sidx / kStSlots) & 1u;
          const int8_t* src = p.x + static_cast<int64_t>(2u * pr + sub) * p.C * p.H * p.in_pitch +
                              static_cast<int64_t>(2 * yc) * p.in_pitch;
          uint4 lo[kUnits], hi[kUnits];
          uint32_t pv[kUnits];                        // the four input bytes left of the unit (its last two feed the shifted copies)
#pragma unroll
          for (int k = 0; k < kUnits; ++k) {
            const int r = 2 * yc + ukh[k] - 3;
            lo[k] = make_uint4(0u, 0u, 0u, 0u); hi[k] = lo[k]; pv[k] = 0u;
            if (r >= 0 && r < p.H && !(p.dbg & 4)) {
              lo[k] = ldg128(src + goff[k]);
              hi[k] = ldg128(src + goff[k] + 16);
              if (!(ufirst[k] & 1)) pv[k] = *reinterpret_cast<const uint32_t*>(src + goff[k] - 4);
            }
          }
          mbar_wait(&a_empty[as], aph ^ 1u);
          uint8_t* dst = smem + kStSmemBar + kStWBytes + as * kStStageBytes;
#pragma unroll
          for (int k = 0; k < kUnits; ++k) {
            if (ukh[k] < 0) continue;
            const uint4 ev = make_uint4(__byte_perm(lo[k].x, lo[k].y, 0x6420), __byte_perm(lo[k].z, lo[k].w, 0x6420),
                                        __byte_perm(hi[k].x, hi[k].y, 0x6420), __byte_perm(hi[k].z, hi[k].w, 0x6420));
            const uint4 od = make_uint4(__byte_perm(lo[k].x, lo[k].y, 0x7531), __byte_perm(lo[k].z, lo[k].w, 0x7531),
                                        __byte_perm(hi[k].x, hi[k].y, 0x7531), __byte_perm(hi[k].z, hi[k].w, 0x7531));
            // shifted right by one pixel: byte 0 comes from the unit on the left (zero at the image edge)
            const uint4 ev1 = make_uint4(__byte_perm(pv[k], ev.x, 0x6542), __funnelshift_l(ev.x, ev.y, 8), __funnelshift_l(ev.y, ev.z, 8),
                                         __funnelshift_l(ev.z, ev.w, 8));
            const uint4 od1 = make_uint4(__byte_perm(pv[k], od.x, 0x6543), __funnelshift_l(od.x, od.y, 8), __funnelshift_l(od.y, od.z, 8),
                                         __funnelshift_l(od.z, od.w, 8));
            *reinterpret_cast<uint4*>(dst + soff[k]) = ev;
            *reinterpret_cast<uint4*>(dst + soff[k] + 4096) = od;
            *reinterpret_cast<uint4*>(dst + soff[k] + 8192) = ev1;
            *reinterpret_cast<uint4*>(dst + soff[k] + 12288) = od1;
            if (ufirst[k] & 2) {                       // last unit of the row: pixel 16 * upr of the shifted copies
              *reinterpret_cast<uint4*>(dst + stail[k] + 8192) = make_uint4(ev.w >> 24, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(dst + stail[k] + 12288) = make_uint4(od.w >> 24, 0u, 0u, 0u);
            }
          }
          fence_proxy_async_smem();                  // generic-proxy stores -> visible to tcgen05.mma
          mbar_arrive(&a_full[as]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kStWarpIssue) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
}

}  // namespace accel
