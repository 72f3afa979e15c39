// Persistent variant of the tcgen05 BSR kernel (bsr_tc.cuh) for activations that arrive by TMA.
//
// One CTA per SM owns all 512 TMEM columns and walks a static list of work items (128-row tile x block-row
// group).  Every role keeps running across item boundaries, so the per-tile costs of the one-shot kernel -
// TMEM allocation, barrier set-up, the first load's latency, the accumulator drain and the epilogue - overlap
// with the next tile's main loop instead of serialising with it:
//
//     TMEM columns [0, 352)     two accumulator sets (11 block-rows x 16 columns each): the MMAs of item n+1
//                               fill one set while the epilogue warps drain (and re-zero) the other
//     TMEM columns [352, 496)   four activation stages of 9 K-tiles; producer half h owns stages h and h+2, so
//                               it re-strides its next stage while the tensor core still reads the previous one
//
//   warps 0-7    activation producers (two halves, alternating stages): ring slot -> 14-in-16 tiles -> tcgen05.st
//   warps 8-15   epilogue: tcgen05.ld, fused requant / ReLU / residual / saturation count, stores, re-zero
//   warps 16-19  MMA issuers (one elected thread each, uniform datapath, schedule from the parameter bank)
//   warp 20      weight loader (cp.async.bulk of B-tile batches)      warp 21   activation loader (TMA tensor tiles)
#pragma once
#include "bsr_tc.cuh"

namespace accel {

constexpr int kPProducerWarps = 8;
constexpr int kPEpilogueWarps = 8;
constexpr int kPWarpIssue = kPProducerWarps + kPEpilogueWarps;     // 16
constexpr int kPWarpWLoad = kPWarpIssue + kIssuers;                // 20
constexpr int kPWarpALoad = kPWarpWLoad + 1;                       // 21
constexpr int kPThreads = (kPWarpALoad + 1) * 32;                  // 704
constexpr int kPTmemCols = 512;
constexpr int kPAccSets = 2;
constexpr int kPXStages = 4;
static_assert(kPAccSets * kAccCols + kPXStages * kXStageCols <= kPTmemCols, "TMEM budget");
constexpr int kPMaxRingSlots = 4;
// shared-memory map
constexpr int kPSmemW = 0;
constexpr int kPSmemBar = kPSmemW + kWStages * kWStageBytes;       // barriers (512 B)
constexpr int kPTabGroups = 8;                                      // block-row groups whose constants stay resident
constexpr int kPSmemScale = kPSmemBar + 512;                       // float [kPTabGroups][176]
constexpr int kPSmemBias = kPSmemScale + kPTabGroups * kAccCols * 4;   // int32 [kPTabGroups][176]
constexpr int kPSmemRing = (kPSmemBias + kPTabGroups * kAccCols * 4 + 127) / 128 * 128;

// ---- int8 epilogue of one block-row for the persistent kernel: 32-bit element offsets (checked on the host),
// residual bytes prefetched by the caller, and an unpredicated variant for the common case "all 14 channels
// exist and all 32 rows of the warp exist".
template <bool RES, bool RES_FAST, bool SAT, bool FULL>
__device__ __forceinline__ void epi_i8_row(const TcParams& p, EpiCtx& ec, const uint32_t (&v)[16], int n_ok, const float* sc,
                                           const int32_t* bi, int8_t* o, int32_t cs, const int (&rv)[kBlock]) {
  if (!FULL && !ec.row_ok) {
    return;
  }
#pragma unroll
  for (int h = 0; h < kBlock; ++h) {
    const int acc = max(static_cast<int>(v[h]) + bi[h], ec.relu_lo);
    const float f = __fmul_rn(__int2float_rn(acc), sc[h]);           // golden_models.cpp:378-411, per channel
    int r8 = cvt_sat_s8(f);
    if constexpr (SAT) {
      // round-half-even leaves [-128, 127] exactly when f >= 127.5 (-> 128) or f < -128.5 (-128.5 -> -128 stays)
      const bool counted = FULL || h < n_ok;
      ec.sat += (counted && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
    }
    if constexpr (RES) {                                             // add_residual_int8, golden_models.cpp:465-490
      const float a = __fmul_rn(__int2float_rn(r8), p.epi.res_scale_main);
      const float r = __fmul_rn(__int2float_rn(rv[h]), p.epi.res_scale_res);
      const float s = __fadd_rn(a, r);
      float d;
      if constexpr (RES_FAST) {    // correctly rounded for every (int8, int8) pair: verified on the host
        const float q0 = __fmul_rn(s, p.res_rcp);
        const float e = __fmaf_rn(-q0, p.epi.res_scale_out, s);
        d = __fmaf_rn(e, p.res_rcp, q0);
      } else {
        d = __fdiv_rn(s, p.epi.res_scale_out);
      }
      r8 = cvt_sat_s8(d);
    }
    const int q = max(r8, ec.out_lo);                                // relu_int8 (golden_models.cpp:278-283) when requested
    if (FULL || h < n_ok) o[h * cs] = static_cast<int8_t>(q);
  }
}
template <bool FULL>
__device__ __forceinline__ void epi_load_residual(const int8_t* __restrict__ r, int32_t cs, bool row_ok, int n_ok,
                                                  int (&rv)[kBlock]) {
#pragma unroll
  for (int h = 0; h < kBlock; ++h) rv[h] = (FULL || (row_ok && h < n_ok)) ? static_cast<int>(r[h * cs]) : 0;
}

// One work item's int8 epilogue for one warp: block-rows g_first, g_first + 2, ...  RESMODE 0 = no residual,
// 1 = residual with the exact 3-instruction divide, 2 = residual with the IEEE divide.  The residual bytes of a
// block-row are requested before the previous one is processed (the first: before the MMAs have even finished).
struct EpiItem {
  uint32_t acc0, g_first, g_rows, g_br0, parity;
  uint64_t* bar;
  const float* s_scale;
  const int32_t* s_bias;
};
template <int RESMODE, bool SAT>
__device__ __forceinline__ void epi_i8_item(const TcParams& p, EpiCtx& ec, const EpiItem& ei) {
  constexpr bool RES = RESMODE != 0;
  const bool rows_all = __all_sync(0xffffffffu, ec.row_ok);
  const int32_t cs32 = static_cast<int32_t>(ec.cs);
  int rv[kBlock];
#pragma unroll
  for (int h = 0; h < kBlock; ++h) rv[h] = 0;
  auto residual = [&](uint32_t g, int (&dst)[kBlock]) {
    const int cb = (ei.g_br0 + g) * kBlock;
    const int n_ok = min(kBlock, p.epi.n_channels - cb);
    const int8_t* r = p.epi.residual + ec.out_base + static_cast<int64_t>(cb) * ec.cs;
    if (rows_all && n_ok == kBlock) epi_load_residual<true>(r, cs32, true, n_ok, dst);
    else epi_load_residual<false>(r, cs32, ec.row_ok, n_ok, dst);
  };
  if (RES && ei.g_first < ei.g_rows) residual(ei.g_first, rv);
  mbar_wait(ei.bar, ei.parity);
  tc_fence_after();
  for (uint32_t g = ei.g_first; g < ei.g_rows; g += 2) {
    uint32_t v[16];
    tmem_ld16(ei.acc0 + g * kTile, v);
    const int cb = (ei.g_br0 + g) * kBlock;
    const int n_ok = min(kBlock, p.epi.n_channels - cb);
    int rvn[kBlock];
#pragma unroll
    for (int h = 0; h < kBlock; ++h) rvn[h] = 0;
    if (RES && g + 2 < ei.g_rows) residual(g + 2, rvn);
    int8_t* o = reinterpret_cast<int8_t*>(p.out) + ec.out_base + static_cast<int64_t>(cb) * ec.cs;
    const float* sc = ei.s_scale + g * kBlock;
    const int32_t* bi = ei.s_bias + g * kBlock;
    tmem_ld_wait();
    if (rows_all && n_ok == kBlock) epi_i8_row<RES, RESMODE == 1, SAT, true>(p, ec, v, n_ok, sc, bi, o, cs32, rv);
    else epi_i8_row<RES, RESMODE == 1, SAT, false>(p, ec, v, n_ok, sc, bi, o, cs32, rv);
    // the accumulators of the next item that uses this set start from zero
#pragma unroll
    for (int c = 0; c < kTile; c += 4) tmem_st4(ei.acc0 + g * kTile + c, 0u, 0u, 0u, 0u);
    if constexpr (RES) {
#pragma unroll
      for (int h = 0; h < kBlock; ++h) rv[h] = rvn[h];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kPThreads, 1) bsr_tcp_kernel(const __grid_constant__ TcLaunch L, uint32_t n_items) {
  static_assert(MODE == kModeGemm || MODE == kModeConv3 || MODE == kModeConv7, "TMA modes only");
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcParams& p = L.p;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPSmemBar);
  uint64_t* x_full = bars;                        // [4]  count 4 (one elected lane per producer warp of the half)
  uint64_t* x_empty = x_full + kPXStages;         // [4]  count kIssuers
  uint64_t* w_full = x_empty + kPXStages;         // [kWStages] count 1 (+tx bytes)
  uint64_t* w_empty = w_full + kWStages;          // [kWStages] count kIssuers
  uint64_t* acc_full = w_empty + kWStages;        // [2]  count kIssuers
  uint64_t* acc_empty = acc_full + kPAccSets;     // [2]  count kPEpilogueWarps
  uint64_t* h_full = acc_empty + kPAccSets;       // [kPMaxRingSlots] count 1 (+tx bytes)
  uint64_t* h_empty = h_full + kPMaxRingSlots;    // [kPMaxRingSlots] count 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_empty + kPMaxRingSlots);
  float* s_scale = reinterpret_cast<float*>(smem + kPSmemScale);
  int32_t* s_bias = reinterpret_cast<int32_t*>(smem + kPSmemBias);

  constexpr bool kHalo = (MODE != kModeGemm);
  constexpr int KS = MODE == kModeConv7 ? 7 : 3;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t n_groups = L.n_groups;
  const uint32_t ring_slots = static_cast<uint32_t>(p.ring_slots);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPXStages; ++s) { mbar_init(&x_full[s], 4); mbar_init(&x_empty[s], kIssuers); }
    for (int s = 0; s < kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], kIssuers); }
    for (int s = 0; s < kPAccSets; ++s) { mbar_init(&acc_full[s], kIssuers); mbar_init(&acc_empty[s], kPEpilogueWarps); }
    for (int s = 0; s < kPMaxRingSlots; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == kPWarpIssue) {
    tmem_alloc_dyn(tmem_slot, kPTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t ring_addr = smem_u32(smem + kPSmemRing);

  if (warp < kPProducerWarps) {
    // =================================================================== activation producers
    const int half = warp >> 2;
    const int tid = threadIdx.x & 127;          // activation row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t t = 0;                             // global stage counter (identical in every role)
    uint32_t slot = 0, sphase = 0;              // ring slot / phase of stage t, advanced without divisions
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
      const uint32_t mt = fdiv(it, p.d_groups);
      const uint32_t gi = it - mt * n_groups;
      const int64_t m0 = static_cast<int64_t>(mt) * kTileM;
      const uint32_t g_bb = L.groups[gi].batch_begin, g_be = L.groups[gi].batch_end;
      uint32_t thr_off = 0, sh8 = 0;
      if constexpr (kHalo) {
        const int64_t m = m0 + tid;
        const uint32_t mm = static_cast<uint32_t>(m < p.M ? m : m0);
        const uint32_t R = fdiv(mm, p.d_wo);
        const int ow = static_cast<int>(mm - R * p.Wo);
        const uint32_t img = fdiv(R, p.d_ho);
        const int oh = static_cast<int>(R - img * p.Ho);
        const uint32_t R0 = fdiv(static_cast<uint32_t>(m0), p.d_wo);
        const uint32_t img0 = fdiv(R0, p.d_ho);
        const uint32_t seg = img - img0;
        const int ohf = seg == 0 ? static_cast<int>(R0 - img0 * p.Ho) : 0;
        const uint32_t xb = static_cast<uint32_t>(ow * p.stride - p.pad + p.halo_lpad);
        thr_off = seg * static_cast<uint32_t>(p.seg_bytes) + static_cast<uint32_t>((oh - ohf) * p.stride) * p.halo_pitch + (xb & ~3u);
        sh8 = (xb & 3u) * 8u;
      }
      for (uint32_t b = g_bb; b < g_be; ++b) {
        const uint32_t bw = L.batches[b];
        if (!(bw & kBatchFirst)) continue;        // one activation stage per K chunk
        const uint32_t a = t & 3u, use = t >> 2, my_slot = slot, my_phase = sphase;
        const bool my = (t & 1u) == static_cast<uint32_t>(half);
        ++t;
        if (++slot == ring_slots) { slot = 0; sphase ^= 1u; }
        if (!my) continue;
        const int chunk = static_cast<int>(bw >> 16);
        const uint32_t xcol = tmem_base + lane_base + kPAccSets * kAccCols + a * kXStageCols;
        mbar_wait(&h_full[my_slot], my_phase);
        mbar_wait(&x_empty[a], (use & 1u) ^ 1u);
        tc_fence_after();
        if constexpr (MODE == kModeGemm) {
          const uint32_t src = ring_addr + my_slot * p.slot_bytes + tid * 144;
          uint32_t r[37];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const uint4 v = lds128(src + i * 16);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
          }
          r[36] = 0u;
          switch ((chunk * kChunkTiles * kBlock) & 15) {
            case 0: gemm_restride<0>(r, xcol); break;
            case 2: gemm_restride<2>(r, xcol); break;
            case 4: gemm_restride<4>(r, xcol); break;
            case 6: gemm_restride<6>(r, xcol); break;
            case 8: gemm_restride<8>(r, xcol); break;
            case 10: gemm_restride<10>(r, xcol); break;
            case 12: gemm_restride<12>(r, xcol); break;
            default: gemm_restride<14>(r, xcol); break;
          }
        } else {
          constexpr int GPS = 126 / KS;
          const uint32_t cs = static_cast<uint32_t>(p.halo_rows) * p.halo_pitch;
          const int g0 = chunk * GPS;
          const int c_first = g0 / KS;
          int kh = g0 - c_first * KS;
          const int groups_left = p.C * KS - g0;
          uint32_t rp = ring_addr + my_slot * p.slot_bytes + thr_off + kh * p.halo_pitch;
          uint32_t w[36];
#pragma unroll
          for (int i = 0; i < 36; ++i) w[i] = 0u;
          ConvBatch<KS>::template run<0>(w, xcol, rp, sh8, p.halo_pitch, cs, kh, groups_left);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&h_empty[my_slot]);
          mbar_arrive(&x_full[a]);
        }
      }
    }
  } else if (warp < kPWarpIssue) {
    // =================================================================== epilogue
    const int ew = warp - kPProducerWarps;            // 0..7
    const int ehalf = ew >> 2;                        // splits the block-rows of a tile between two warp sets
    const int etid = threadIdx.x - kPProducerWarps * 32;   // 0..255
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    // both accumulator sets start out zero
    for (uint32_t c = ehalf * 4u; c < static_cast<uint32_t>(kPAccSets * kAccCols); c += 8)
      tmem_st4(tmem_base + lane_base + c, 0u, 0u, 0u, 0u);
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1, kPEpilogueWarps * 32);
    if (lane == 0) { mbar_arrive(&acc_empty[0]); mbar_arrive(&acc_empty[1]); }
    EpiCtx ec;
    ec.flags = p.epi.flags;
    ec.cs = p.lay.chan_stride;
    ec.relu_lo = (ec.flags & ACCEL_RELU) ? 0 : INT_MIN;
    ec.out_lo = (ec.flags & ACCEL_RELU_OUT) ? 0 : -128;
    ec.sat = 0;
    ec.lane = lane;
    const bool sat_on = p.epi.sat_count != nullptr;
    int kind = kEpiGeneric;
    if (!p.epi.chan_absmax && (ec.flags & ACCEL_OUT_I8))
      kind = !p.epi.residual ? kEpiI8 : (p.res_fast ? kEpiI8ResFast : kEpiI8Res);
    else if (!p.epi.chan_absmax && !sat_on && (ec.flags & ACCEL_OUT_I32)) kind = kEpiI32;
    if (sat_on && kind != kEpiGeneric) kind += 8;
    const bool tabs_resident = n_groups <= static_cast<uint32_t>(kPTabGroups);
    if (tabs_resident) {
      for (uint32_t i = etid; i < n_groups * kAccCols; i += kPEpilogueWarps * 32) {
        const uint32_t g = i / kAccCols, j = i - g * kAccCols;
        const uint32_t br0 = L.groups[g].br0_rows & 0xffffu, rows = L.groups[g].br0_rows >> 16;
        const int c = br0 * kBlock + j;
        const bool ok = j < rows * kBlock && c < p.epi.n_channels;
        s_scale[i] = (ok && p.epi.chan_scale) ? p.epi.chan_scale[c] : 0.f;
        s_bias[i] = (ok && p.epi.bias) ? p.epi.bias[c] : 0;
      }
      named_bar_sync(1, kPEpilogueWarps * 32);
    }
    uint32_t n = 0;
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
      const uint32_t mt = fdiv(it, p.d_groups);
      const uint32_t gi = it - mt * n_groups;
      const int64_t m0 = static_cast<int64_t>(mt) * kTileM;
      const uint32_t g_br0 = L.groups[gi].br0_rows & 0xffffu, g_rows = L.groups[gi].br0_rows >> 16;
      const uint32_t ab = n & 1u;
      // per-channel constants: resident for every group of the launch when there are few groups (loaded once,
      // above); otherwise reloaded per item into table 0 (the previous item's readers are done: barrier first)
      const uint32_t tab = tabs_resident ? gi : 0u;
      if (!tabs_resident) {
        named_bar_sync(1, kPEpilogueWarps * 32);
        if (static_cast<uint32_t>(etid) < g_rows * kBlock) {
          const int c = g_br0 * kBlock + etid;
          const bool ok = c < p.epi.n_channels;
          s_scale[etid] = (ok && p.epi.chan_scale) ? p.epi.chan_scale[c] : 0.f;
          s_bias[etid] = (ok && p.epi.bias) ? p.epi.bias[c] : 0;
        }
      }
      const int64_t m = m0 + (warp & 3) * 32 + lane;
      ec.row_ok = m < p.M;
      ec.out_base = 0;
      if (ec.row_ok) {
        // M < 2^31 (checked by the host) and rows_per_image < 2^31 whenever d_rpi.d is its divisor
        const uint32_t mu = static_cast<uint32_t>(m);
        const uint32_t im = p.lay.rows_per_image < (1ll << 31) ? fdiv(mu, p.d_rpi) : 0u;
        const uint32_t pix = mu - im * static_cast<uint32_t>(p.lay.rows_per_image);
        if (p.lay.row_len > 0) {
          const uint32_t r = fdiv(pix, p.d_rowlen);
          ec.out_base = static_cast<int64_t>(im) * p.lay.image_stride + static_cast<int64_t>(r) * p.lay.row_pitch +
                        static_cast<int64_t>(pix - r * static_cast<uint32_t>(p.lay.row_len)) * p.lay.row_stride;
        } else {
          ec.out_base = static_cast<int64_t>(im) * p.lay.image_stride + static_cast<int64_t>(pix) * p.lay.row_stride;
        }
      }
      if (!tabs_resident) named_bar_sync(1, kPEpilogueWarps * 32);
      const bool fast8 = p.out_small && (kind & 7) <= kEpiI8Res;      // int8 output with 32-bit element offsets
      if (fast8) {
        EpiItem ei;
        ei.acc0 = tmem_base + lane_base + ab * kAccCols;
        ei.g_first = static_cast<uint32_t>(ehalf); ei.g_rows = g_rows; ei.g_br0 = g_br0;
        ei.bar = &acc_full[ab]; ei.parity = (n >> 1) & 1u;
        ei.s_scale = s_scale + tab * kAccCols; ei.s_bias = s_bias + tab * kAccCols;
        switch ((kind & 7) * 2 + (sat_on ? 1 : 0)) {
          case 0: epi_i8_item<0, false>(p, ec, ei); break;
          case 1: epi_i8_item<0, true>(p, ec, ei); break;
          case 2: epi_i8_item<1, false>(p, ec, ei); break;
          case 3: epi_i8_item<1, true>(p, ec, ei); break;
          case 4: epi_i8_item<2, false>(p, ec, ei); break;
          default: epi_i8_item<2, true>(p, ec, ei); break;
        }
      } else {
        mbar_wait(&acc_full[ab], (n >> 1) & 1u);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + lane_base + ab * kAccCols;
        for (uint32_t g = ehalf; g < g_rows; g += 2) {
          uint32_t v[16];
          tmem_ld16(acc0 + g * kTile, v);
          const int cb = (g_br0 + g) * kBlock;
          const int n_ok = min(kBlock, p.epi.n_channels - cb);
          const float* sc = s_scale + tab * kAccCols + g * kBlock;
          const int32_t* bi = s_bias + tab * kAccCols + g * kBlock;
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
          // the accumulators of the next item that uses this set start from zero
#pragma unroll
          for (int c = 0; c < kTile; c += 4) tmem_st4(acc0 + g * kTile + c, 0u, 0u, 0u, 0u);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
    if (p.epi.sat_count) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, ec.sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
  } else if (warp < kPWarpWLoad) {
    // =================================================================== MMA issuers (uniform datapath)
    const uint32_t me = static_cast<uint32_t>(warp - kPWarpIssue);
    if (elect_one()) {
      const uint32_t idesc0 = idesc_i8(kTileM, 0);
      const uint32_t w_addr = smem_u32(smem + kPSmemW);
      const uint64_t bdesc0 = smem_desc_kmajor(0, 128, 256);
      const uint32_t bdesc_hi = static_cast<uint32_t>(bdesc0 >> 32);
      const uint32_t bdesc_lo0 = static_cast<uint32_t>(bdesc0);
      uint32_t t = 0, wcount = 0, n = 0, a = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const uint32_t gi = it - fdiv(it, p.d_groups) * n_groups;
        const uint32_t g_bb = L.groups[gi].batch_begin, g_be = L.groups[gi].batch_end;
        uint32_t opb = L.groups[gi].op_begin;
        const uint32_t ab = n & 1u;
        mbar_wait(&acc_empty[ab], (n >> 1) & 1u);      // drained and re-zeroed by the epilogue warps
        tc_fence_after();
        const uint32_t accb = tmem_base + ab * kAccCols;
        for (uint32_t b = g_bb; b < g_be; ++b) {
          const uint32_t bw = L.batches[b];
          if (bw & kBatchFirst) {
            a = t & 3u;
            mbar_wait(&x_full[a], (t >> 2) & 1u);
            ++t;
          }
          const uint32_t ws_i = wcount % kWStages;
          mbar_wait(&w_full[ws_i], (wcount / kWStages) & 1u);
          tc_fence_after();
          const uint32_t n_runs = bw & 0xffu;
          const uint32_t xa = tmem_base + kPAccSets * kAccCols + a * kXStageCols;
          const uint32_t blo = bdesc_lo0 | (((w_addr + ws_i * kWStageBytes) >> 4) & 0x3FFFu);
          for (uint32_t i = me; i < n_runs; i += kIssuers) {
            const OpRec o = L.ops[opb + i];
            const uint32_t d = accb + (o.d_n & 0x1ffu);
            const uint32_t idesc = idesc0 | (o.d_n & 0x7E0000u);
            const uint32_t aa = xa + (o.a_b & 0x1ffu);
            const uint64_t bdesc = (static_cast<uint64_t>(bdesc_hi) << 32) | (blo + (o.a_b >> 16));
            mma_i8_ts(d, aa, bdesc, idesc, 1u);
          }
          opb += n_runs;
          mma_commit(&w_empty[ws_i]);
          if (bw & kBatchLast) mma_commit(&x_empty[a]);
          ++wcount;
        }
        mma_commit(&acc_full[ab]);    // with nothing in flight (a group without blocks) this arrives at once
      }
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp == kPWarpWLoad) {
    // =================================================================== weight loader
    if (lane == 0) {
      uint32_t wcount = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t gi = it - fdiv(it, p.d_groups) * n_groups;
        const uint32_t g_bb = L.groups[gi].batch_begin, g_be = L.groups[gi].batch_end;
        const uint8_t* src = p.blob + static_cast<size_t>(L.groups[gi].blob_off16) * 16;
        for (uint32_t b = g_bb; b < g_be; ++b) {
          const uint32_t bw = L.batches[b];
          const uint32_t ws_i = wcount % kWStages;
          mbar_wait(&w_empty[ws_i], ((wcount / kWStages) & 1u) ^ 1u);
          const uint32_t bytes = (((bw >> 8) & 0x3fu) + 1u) * kBTileBytes;
          mbar_arrive_expect_tx(&w_full[ws_i], bytes);
          bulk_g2s(smem + kPSmemW + ws_i * kWStageBytes, src, bytes, &w_full[ws_i]);
          src += bytes;
          ++wcount;
        }
      }
    }
    __syncwarp();
  } else {
    // =================================================================== activation loader (TMA)
    if (elect_one()) {
      uint32_t slot = 0, sphase = 0;
      for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const uint32_t mt = fdiv(it, p.d_groups);
        const uint32_t gi = it - mt * n_groups;
        const int64_t m0 = static_cast<int64_t>(mt) * kTileM;
        const uint32_t g_bb = L.groups[gi].batch_begin, g_be = L.groups[gi].batch_end;
        uint32_t n_seg = 1, img0 = 0;
        int ih_first = 0;
        if constexpr (kHalo) {
          const uint32_t R0 = fdiv(static_cast<uint32_t>(m0), p.d_wo);
          img0 = fdiv(R0, p.d_ho);
          const uint32_t m_last = static_cast<uint32_t>(m0 + kTileM - 1 < p.M ? m0 + kTileM - 1 : p.M - 1);
          n_seg = m_last / static_cast<uint32_t>(p.Ho * p.Wo) - img0 + 1;
          ih_first = static_cast<int>(R0 - img0 * p.Ho) * p.stride - p.pad;
        }
        for (uint32_t b = g_bb; b < g_be; ++b) {
          const uint32_t bw = L.batches[b];
          if (!(bw & kBatchFirst)) continue;
          mbar_wait(&h_empty[slot], sphase ^ 1u);
          tma_stage<MODE>(L, ring_addr + slot * p.slot_bytes, &h_full[slot], static_cast<int>(bw >> 16), m0, n_seg, img0,
                          ih_first);
          if (++slot == ring_slots) { slot = 0; sphase ^= 1u; }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kPWarpIssue) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, kPTmemCols);
  }
}

}  // namespace accel
