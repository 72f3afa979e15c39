// Dense-equivalent BSR GEMM on tcgen05 (sm_100a), CTA pairs: the kernel of the 4096^3 sweep, the FC layers and every
// [M, K] x BSR product with 16-byte aligned activation rows (golden_fc1_test.py:49-108 is what it computes).
//
//   D[channel][row] += Wd[channel][128 k] x X[row][128 k]^T          TMEM lane = output channel, TMEM column = activation row
//
//   A operand  the BSR weights, scattered once (accel_plan_gemm_ws_prepare) from the stored 14x14 blocks into a plain
//              row-major int8 matrix Wd[channels, K] (blocks that are not stored stay zero).  Neither the channels nor K
//              are padded 14 -> 16: block (br, bc) lands at rows 14 br, columns 14 bc, so every MMA contracts 32 real K
//              values for 128 real channels.  A (128 * CG channels) x (128 k) region without any non-zero weight is never
//              loaded nor multiplied: the issuer and the loader walk a per-channel-tile list of live K chunks built from
//              the data at prepare time.
//   B operand  the activation matrix X[M, K] itself (row-major int8, rows 16-byte aligned), 256 rows per tile.
//   both       arrive by TMA as 128-row x 128-byte boxes with the 128-byte swizzle, K-major; one stage = 128 k.
//
// CG = 2 (default): a CTA pair shares one 256 x 256 tile through tcgen05.mma.cta_group::2 - each CTA loads 128 channels of
// Wd and 128 rows of X, the leader issues M = 256, N = 256 MMAs that read both shared memories: per-SM operand traffic
// from L2 is half that of two independent 128 x 256 tiles.  CG = 1 is the same kernel with 128 x 256 tiles per CTA.
// Persistent: one CTA (pair) per SM (pair), static round-robin over the tiles, two accumulator sets (2 x 256 TMEM
// columns) so that the epilogue of a tile overlaps the MMAs of the next.
// Epilogue: thread = output channel (bias / requant factor in registers), a warp's 32 lanes store 32 consecutive channels
// of one output row: 128 contiguous bytes (INT32 / float) or 32 (INT8) per store instruction.
#pragma once
#include "conv_ws.cuh"

namespace accel {

constexpr int kGwCh = 128;                    // channels per CTA (TMEM lanes)
constexpr int kGwRows = 256;                  // activation rows per tile (MMA N, TMEM columns of one accumulator set)
constexpr int kGwKc = 128;                    // K bytes per stage
constexpr int kGwBoxBytes = 128 * kGwKc;      // one TMA box: 128 rows x 128 bytes
constexpr int kGwEpiWarps = 8;
constexpr int kGwWarpTma = kGwEpiWarps;       // 8
constexpr int kGwWarpMma = kGwEpiWarps + 1;   // 9
constexpr int kGwThreads = (kGwWarpMma + 1) * 32;   // 320
constexpr int kGwMaxStages = 8;
constexpr int kGwSmemBar = 1024;

struct GwParams {
  int64_t M;
  int32_t n_ch_tiles;          // tiles of 128 * CG channels
  int32_t n_row_tiles;         // tiles of 256 rows
  int32_t n_stages;
  int32_t klist_stride;
  const uint16_t* klist;       // [n_ch_tiles][klist_stride]: live K chunks (of 128) of the channel tile, ascending
  const uint16_t* kcount;      // [n_ch_tiles]
  int32_t k_chunks_x;          // K chunks the activation matrix covers: later chunks multiply zeros and are skipped
  accel_epilogue epi;
  int32_t res_fast;
  float res_rcp;
  void* out;
  accel_out_layout lay;
  FastDiv d_rpi;
  int32_t single_image;        // rows_per_image >= M: offset = m * row_stride + c * chan_stride
  int32_t dbg;
};
struct GwLaunch {
  alignas(64) CUtensorMap tmap_w;
  alignas(64) CUtensorMap tmap_x;
  GwParams p;
};

// ---- prepare-time kernels -------------------------------------------------------------------------------------------
// Scatter the stored 14x14 blocks into the dense row-major matrix (zeroed first).
__global__ void gw_scatter_kernel(const int8_t* __restrict__ blocks, const int32_t* __restrict__ blk_row,
                                  const int32_t* __restrict__ col_idx, int64_t nnz, int64_t ld, int8_t* __restrict__ wd) {
  const int64_t b = blockIdx.x;
  if (b >= nnz) return;
  const int64_t r0 = static_cast<int64_t>(blk_row[b]) * kBlock, c0 = static_cast<int64_t>(col_idx[b]) * kBlock;
  for (int i = threadIdx.x; i < kBlock * kBlock; i += blockDim.x) {
    const int h = i / kBlock, w = i - h * kBlock;
    wd[(r0 + h) * ld + c0 + w] = blocks[b * (kBlock * kBlock) + i];
  }
}
// flags[t][j] = 1 when rows [128 t, 128 t + 128), columns [128 j, 128 j + 128) of Wd hold any non-zero byte
__global__ void gw_flags_kernel(const int8_t* __restrict__ wd, int64_t ld, int32_t n_chunks, uint8_t* __restrict__ flags) {
  const int t = blockIdx.y, j = blockIdx.x;
  const int8_t* base = wd + static_cast<int64_t>(t) * 128 * ld + static_cast<int64_t>(j) * 128;
  uint32_t any = 0;
  for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, q = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(base + static_cast<int64_t>(r) * ld + q * 16);
    any |= v.x | v.y | v.z | v.w;
  }
  const int nz = __syncthreads_or(any != 0);
  if (threadIdx.x == 0) flags[t * n_chunks + j] = nz ? 1 : 0;
}

// ---- epilogue arithmetic (golden_models.cpp:298-303, :378-411, :465-490) ---------------------------------
struct GwEpiConst {
  int bias, relu_lo, out_lo;
  float sf;
};
__device__ __forceinline__ int gw_requant(int acc, const GwEpiConst& k, uint32_t& sat, bool count) {
  const float f = __fmul_rn(__int2float_rn(acc), k.sf);
  sat += (count && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
  int q;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(q) : "f"(f));
  return static_cast<int>(static_cast<int8_t>(q));       // only the low byte of the destination register is defined
}
__device__ __forceinline__ int gw_residual(int q, int r, const GwParams& p) {
  const float a = __fmul_rn(__int2float_rn(q), p.epi.res_scale_main);
  const float b = __fmul_rn(__int2float_rn(r), p.epi.res_scale_res);
  const float sm = __fadd_rn(a, b);
  float d;
  if (p.res_fast) {      // exact for every (int8, int8) pair: verified on the host (residual_divide_mode)
    const float q0 = __fmul_rn(sm, p.res_rcp);
    const float er = __fmaf_rn(-q0, p.epi.res_scale_out, sm);
    d = __fmaf_rn(er, p.res_rcp, q0);
  } else {
    d = __fdiv_rn(sm, p.epi.res_scale_out);
  }
  int o;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(o) : "f"(d));
  return static_cast<int>(static_cast<int8_t>(o));
}

template <int OUTK>
__device__ __forceinline__ void gw_store_one(const GwParams& p, uint32_t v, int64_t off, const GwEpiConst& kc, uint32_t& sat, bool count) {
  const int a = max(static_cast<int>(v) + kc.bias, kc.relu_lo);
  if constexpr (OUTK == 0) {
    static_cast<int32_t*>(p.out)[off] = a;
  } else if constexpr (OUTK == 2) {
    static_cast<float*>(p.out)[off] = __fmul_rn(__int2float_rn(a), kc.sf);
  } else {
    int q = gw_requant(a, kc, sat, count);
    if (p.epi.residual) q = gw_residual(q, static_cast<int>(__ldg(p.epi.residual + off)), p);
    static_cast<int8_t*>(p.out)[off] = static_cast<int8_t>(max(q, kc.out_lo));
  }
}
// Store up to 32 consecutive activation rows of this thread's channel: element offsets off0 + j * rs.
template <int OUTK, bool FULL>
__device__ __forceinline__ void gw_store_rows(const GwParams& p, const uint32_t (&v)[32], int64_t off0, int64_t rs, int n_valid,
                                              const GwEpiConst& kc, uint32_t& sat, bool count) {
  if constexpr (OUTK == 0) {
    int32_t* o = static_cast<int32_t*>(p.out) + off0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (FULL || j < n_valid) o[j * rs] = max(static_cast<int>(v[j]) + kc.bias, kc.relu_lo);
  } else if constexpr (OUTK == 2) {
    float* o = static_cast<float*>(p.out) + off0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (FULL || j < n_valid) o[j * rs] = __fmul_rn(__int2float_rn(max(static_cast<int>(v[j]) + kc.bias, kc.relu_lo)), kc.sf);
  } else {
    int8_t* o = static_cast<int8_t*>(p.out) + off0;
    const int8_t* r = p.epi.residual ? p.epi.residual + off0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (FULL || j < n_valid) {
        int q = gw_requant(max(static_cast<int>(v[j]) + kc.bias, kc.relu_lo), kc, sat, count);
        if (r) q = gw_residual(q, static_cast<int>(__ldg(r + j * rs)), p);
        o[j * rs] = static_cast<int8_t>(max(q, kc.out_lo));
      }
  }
}

// OUTK: 0 = INT32 accumulators, 1 = INT8 (requant, optional residual), 2 = float32 (de-quantised)
template <int CG, int OUTK>
__global__ void __launch_bounds__(kGwThreads, 1) gemm_ws_kernel(const __grid_constant__ GwLaunch L) {
  extern __shared__ uint8_t smem_dyn[];
  const GwParams& p = L.p;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;     // identical in every CTA: same static layout
  uint8_t* smem = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full = bars;                               // [kGwMaxStages]
  uint64_t* empty = full + kGwMaxStages;               // [kGwMaxStages]
  uint64_t* acc_full = empty + kGwMaxStages;           // [2]
  uint64_t* acc_empty = acc_full + 2;                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  constexpr uint32_t kXBytes = CG == 2 ? kGwBoxBytes : 2 * kGwBoxBytes;      // activation rows this CTA stages: 128 / 256
  constexpr uint32_t kStageBytes = kGwBoxBytes + kXBytes;
  const uint32_t stage0 = base + kGwSmemBar;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const uint32_t unit = CG == 2 ? blockIdx.x >> 1 : blockIdx.x;           // CTA (pair) index
  const uint32_t n_units = CG == 2 ? gridDim.x >> 1 : gridDim.x;
  const uint32_t n_ct = static_cast<uint32_t>(p.n_ch_tiles);
  const uint32_t n_tiles = n_ct * static_cast<uint32_t>(p.n_row_tiles);
  const uint32_t n_stages = static_cast<uint32_t>(p.n_stages);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGwMaxStages; ++s) { mbar_init(&full[s], CG); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kGwEpiWarps * CG); }
    fence_mbar_init();
  }
  if (warp == kGwWarpMma) {
    tmem_alloc_cg<CG>(tmem_slot, 512);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kGwEpiWarps) {
    // =================================================================== epilogue: thread = output channel
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    GwEpiConst kc;
    kc.relu_lo = (p.epi.flags & ACCEL_RELU) ? 0 : INT_MIN;
    kc.out_lo = (p.epi.flags & ACCEL_RELU_OUT) ? 0 : -128;
    const bool count = p.epi.sat_count != nullptr;
    uint32_t sat = 0, n = 0;
    const uint32_t empty_addr = CG == 2 ? mapa_u32(smem_u32(&acc_empty[0]), 0u) : smem_u32(&acc_empty[0]);
    for (uint32_t t = unit; t < n_tiles; t += n_units, ++n) {
      const uint32_t rt = t / n_ct, ct = t - rt * n_ct;
      const int co = static_cast<int>(ct * (kGwCh * CG) + rank * kGwCh) + q * 32 + lane;
      const bool ch_ok = co < p.epi.n_channels;
      kc.sf = (ch_ok && p.epi.chan_scale) ? p.epi.chan_scale[co] : 0.f;
      kc.bias = (ch_ok && p.epi.bias) ? p.epi.bias[co] : 0;
      // no live chunk inside the activation matrix's K range: the accumulators were never written
      const bool live = p.kcount[ct] != 0 && static_cast<int>(p.klist[static_cast<size_t>(ct) * p.klist_stride]) < p.k_chunks_x;
      const int64_t cbase = static_cast<int64_t>(co) * p.lay.chan_stride;
      const int64_t m0 = static_cast<int64_t>(rt) * kGwRows + half * 128;
      const uint32_t ab = n & 1u;
      mbar_wait(&acc_full[ab], (n >> 1) & 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + ab * kGwRows + half * 128;
      if (!(ACCEL_DEV && (p.dbg & 1)) && m0 < p.M && __any_sync(0xffffffffu, ch_ok)) {
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int64_t mc = m0 + cc * 32;
          if (mc >= p.M) break;
          uint32_t v[32];
          if (live) {
            tmem_ld32(acc + cc * 32, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
          }
          const int n_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - mc));
          // element offset of (row mc, this thread's channel) and the stride between consecutive rows; rows of one chunk that
          // straddle two images of a multi-image layout take the per-element path
          int64_t off0 = cbase, rs = p.lay.row_stride;
          bool chunk_contig = true;
          if (p.single_image) {
            off0 += mc * rs;
          } else {
            const uint32_t img = fdiv(static_cast<uint32_t>(mc), p.d_rpi);
            const int64_t pix0 = mc - static_cast<int64_t>(img) * p.lay.rows_per_image;
            off0 += static_cast<int64_t>(img) * p.lay.image_stride + pix0 * rs;
            chunk_contig = pix0 + n_valid <= p.lay.rows_per_image;
          }
          if (!ch_ok) continue;
          if (chunk_contig) {
            if (n_valid == 32) gw_store_rows<OUTK, true>(p, v, off0, rs, 32, kc, sat, count);
            else gw_store_rows<OUTK, false>(p, v, off0, rs, n_valid, kc, sat, count);
          } else {
#pragma unroll 1
            for (int j = 0; j < n_valid; ++j) {
              const int64_t m = mc + j;
              const uint32_t img = fdiv(static_cast<uint32_t>(m), p.d_rpi);
              const int64_t off = cbase + static_cast<int64_t>(img) * p.lay.image_stride +
                                  (m - static_cast<int64_t>(img) * p.lay.rows_per_image) * rs;
              uint32_t val = v[0];        // rare path (a chunk that crosses an image boundary): select without local memory
#pragma unroll
              for (int q2 = 1; q2 < 32; ++q2) val = q2 == j ? v[q2] : val;
              gw_store_one<OUTK>(p, val, off, kc, sat, count);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(empty_addr + ab * 8u);
        else mbar_arrive(&acc_empty[ab]);
      }
    }
    if (OUTK == 1 && count) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
  } else if (warp == kGwWarpTma) {
    // =================================================================== loader: TMA boxes of Wd and X (both CTAs of a pair)
    if (elect_one()) {
      uint32_t s = 0, ph = 0;
      const uint32_t full0 = CG == 2 ? mapa_u32(smem_u32(&full[0]), 0u) : smem_u32(&full[0]);   // the leader's barriers
      for (uint32_t t = unit; t < n_tiles; t += n_units) {
        const uint32_t rt = t / n_ct, ct = t - rt * n_ct;
        const uint32_t cnt = p.kcount[ct];
        const uint16_t* list = p.klist + static_cast<size_t>(ct) * p.klist_stride;
        const int ch0 = static_cast<int>(ct * (kGwCh * CG) + rank * kGwCh);
        const int m0 = static_cast<int>(rt * kGwRows + (CG == 2 ? rank * 128u : 0u));
        for (uint32_t i = 0; i < cnt; ++i) {
          const uint32_t kc = list[i];
          if (static_cast<int>(kc) >= p.k_chunks_x) break;            // ascending: nothing of X beyond this point
          mbar_wait(&empty[s], ph ^ 1u);
          const uint32_t dst = stage0 + s * kStageBytes;
          const uint32_t bar = full0 + s * 8u;
          if constexpr (CG == 2) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2u * kStageBytes);
            else mbar_arrive_cluster(bar);
          } else {
            mbar_arrive_expect_tx(&full[s], kStageBytes);
          }
          tma_load_2d_cg<CG>(dst, &L.tmap_w, static_cast<int>(kc) * kGwKc, ch0, bar);
          tma_load_2d_cg<CG>(dst + kGwBoxBytes, &L.tmap_x, static_cast<int>(kc) * kGwKc, m0, bar);
          if constexpr (CG == 1) tma_load_2d_cg<CG>(dst + 2 * kGwBoxBytes, &L.tmap_x, static_cast<int>(kc) * kGwKc, m0 + 128, bar);
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // =================================================================== MMA issuer (the leader CTA of a pair)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = idesc_i8(static_cast<uint32_t>(kGwCh * CG), static_cast<uint32_t>(kGwRows));
      const uint64_t desc0 = smem_desc_any(0, 16, 1024, 2u);          // K-major, 128-byte swizzle, 8-row groups 1024 B apart
      const uint32_t d_hi = static_cast<uint32_t>(desc0 >> 32), d_lo0 = static_cast<uint32_t>(desc0);
      uint32_t s = 0, ph = 0, n = 0;
      for (uint32_t t = unit; t < n_tiles; t += n_units, ++n) {
        const uint32_t rt = t / n_ct, ct = t - rt * n_ct;
        uint32_t cnt = p.kcount[ct];
        const uint16_t* list = p.klist + static_cast<size_t>(ct) * p.klist_stride;
        while (cnt > 0 && static_cast<int>(list[cnt - 1]) >= p.k_chunks_x) --cnt;      // the loader stops at the same chunk
        const uint32_t ab = n & 1u;
        mbar_wait(&acc_empty[ab], ((n >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + ab * kGwRows;
        for (uint32_t i = 0; i < cnt; ++i) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t wl = d_lo0 | (((stage0 + s * kStageBytes) >> 4) & 0x3FFFu);
          const uint32_t xl = d_lo0 | (((stage0 + s * kStageBytes + kGwBoxBytes) >> 4) & 0x3FFFu);
          if (!(ACCEL_DEV && (p.dbg & 2))) {
#pragma unroll
            for (uint32_t k = 0; k < kGwKc / 32; ++k)
              mma_i8_ss_cg<CG>(d, (static_cast<uint64_t>(d_hi) << 32) | (wl + 2u * k), (static_cast<uint64_t>(d_hi) << 32) | (xl + 2u * k),
                               idesc, (i | k) != 0u ? 1u : 0u);
          }
          mma_commit_cg<CG>(&empty[s]);
          if (++s == n_stages) { s = 0; ph ^= 1u; }
        }
        mma_commit_cg<CG>(&acc_full[ab]);          // with nothing in flight this arrives at once
      }
    }
    __syncwarp();
    tc_fence_before();
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kGwWarpMma) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

}  // namespace accel
