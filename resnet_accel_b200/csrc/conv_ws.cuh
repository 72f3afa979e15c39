// Weight-stationary 3x3 / stride-1 / pad-1 BSR convolution on tcgen05 (sm_100a): no register-path im2col at all.
//
//   D[co][px]  +=  W_tap[co][32 ch]  x  X[32 ch][px]          TMEM lane = output channel, TMEM column = pixel
//
//   A operand  the BSR weights of the layer, scattered once (accel_plan_conv_ws_prepare) from the stored 14x14
//              blocks into K-major 128 x 32 tiles  [channel group of 128][chunk of 32 input channels][tap (kh,kw)];
//              blocks that are not stored stay zero, (group, chunk, tap) tiles without any stored weight are
//              skipped by the issuer (9-bit masks built on the host from row_ptr / col_idx).
//   B operand  the activation tile itself: ONE TMA tensor tile [32 ch][R+2 input rows][P pixels] per chunk, taken
//              from the NCHW tensor with the dimensions ordered (x, c, y, n), lands in shared memory as an MN-major
//              (pixel-contiguous) operand whose swizzle atom is one image row of 8 channels.  Rows / columns outside
//              the image are zero-filled by TMA - that is the convolution padding.
//   tap (kh, kw)
//              kh moves the B window by one image row (descriptor start + kh * row_stride): all nine taps read
//              the same staged tile.  kw is a shift by one pixel = one TMEM column, which neither TMA (16-byte
//              aligned coordinates only) nor the accumulator address (even columns only) can express directly:
//              kw = 1 accumulates into Z1 (columns [0, N)), kw = 0 and kw = 2 into U with a relative shift of two
//              columns (kw = 0 at U + 2, kw = 2 at U + 0), and the epilogue reads  Y[p] = Z1[p] + U[p + 1]
//              (tcgen05.ld accepts odd columns).  U = Z1 + N - 2: its first two columns alias the last two of
//              Z1, which belong to padding pixels of the tile's last row and receive exact zeros.
//              Measured (tools/probe/mma_probe9): bit-exact, 66.8 cycles per 128x128x32 MMA (floor 64).
//
// One persistent CTA per SM: 8 epilogue warps (thread = output channel, 16 consecutive pixels per tcgen05.ld:
// per-channel constants live in registers, residual loads and output stores are 16-byte vectors), one MMA issuer,
// one loader (TMA activation tiles + bulk copies of the weight tiles; the weights of a channel group stay resident
// in shared memory when they fit, else they stream through a ring).  Two accumulator sets (2 x 256 TMEM columns).
#pragma once
#include "bsr_tc.cuh"

namespace accel {

constexpr int kWsEpiWarps = 8;
constexpr int kWsWarpIssue = kWsEpiWarps;          // 8
constexpr int kWsWarpLoad = kWsWarpIssue + 1;      // 9
constexpr int kWsThreads = (kWsWarpLoad + 1) * 32; // 320
constexpr int kWsCo = 128;                         // output channels per group (TMEM lanes)
constexpr int kWsCk = 32;                          // input channels per chunk (one MMA K)
constexpr int kWsTapBytes = kWsCo * kWsCk;         // 4096
constexpr int kWsChunkBytes = 9 * kWsTapBytes;     // 36864
constexpr int kWsMaxChunks = 16;                   // Cin <= 512
constexpr int kWsMaxGroups = 8;                    // Cout <= 1024
constexpr int kWsMaxWSlots = 4;                    // resident: Cin <= 128; else a ring of this many chunk slots
constexpr int kWsMaxASlots = 8;
constexpr int kWsAccCols = 256;                    // per accumulator set: Z1 [0, N), U [N-2, 2N)
constexpr int kWsSmemBar = 1024;                   // barriers + TMEM slot in front of the operand areas

struct WsParams {
  int32_t C, H, W, B;          // input geometry (output has the same H, W)
  int32_t P, R, N;             // pixels per staged row (16 / 32 / 64), output rows per tile, N = R * P
  int32_t n_chunks, n_groups, c_out;
  int32_t tiles_per_image, n_tiles;   // row tiles per image, B * tiles_per_image
  int32_t w_slots, w_resident, a_slots, a_stage_bytes, a_box_bytes;
  uint32_t b_layout, b_lbo, b_sbo, row_stride;
  FastDiv d_tpi;
  const uint8_t* wblob;        // [group][chunk][tap][4096]
  accel_epilogue epi;
  int32_t res_fast;
  float res_rcp;
  int8_t* out;
  int32_t out_pitch;           // bytes between output rows
  int32_t chan_stride;         // H * out_pitch
  int32_t x_store_end;         // pixels of a row that are stored (W rounded up to 16)
  int64_t image_stride;        // c_out * chan_stride
  uint16_t masks[kWsMaxGroups * kWsMaxChunks];
};
struct WsLaunch {
  alignas(64) CUtensorMap tmap;
  WsParams p;
};

__host__ __device__ constexpr uint32_t idesc_i8_bmn(uint32_t M, uint32_t N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (0u << 15) /* A K-major */ | (1u << 16) /* B MN-major */ | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}
// generic shared-memory matrix descriptor (layout: 0 none, 6 = 32 B, 4 = 64 B, 2 = 128 B swizzle)
__device__ __forceinline__ uint64_t smem_desc_any(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// Scatter the stored 14x14 blocks into the weight-stationary layout (the blob is zeroed first).
// blk_row[b] = block-row of stored block b; K index k = c * 9 + tap (golden_models.cpp:801-842).
__global__ void ws_scatter_kernel(const int8_t* __restrict__ blocks, const int32_t* __restrict__ blk_row,
                                  const int32_t* __restrict__ col_idx, int64_t nnz, int32_t c_in, int32_t c_out, int32_t n_chunks,
                                  uint8_t* __restrict__ blob) {
  const int64_t b = blockIdx.x;
  if (b >= nnz) return;
  const int br = blk_row[b], bc = col_idx[b];
  for (int i = threadIdx.x; i < kBlock * kBlock; i += blockDim.x) {
    const int h = i / kBlock, w = i - h * kBlock;
    const int co = br * kBlock + h, k = bc * kBlock + w;
    if (co >= c_out || k >= c_in * 9) continue;
    const int c = k / 9, tap = k - c * 9;
    const int g = co / kWsCo, row = co - g * kWsCo, j = c / kWsCk, kk = c - j * kWsCk;
    const size_t dst = (static_cast<size_t>(g * n_chunks + j) * 9 + tap) * kWsTapBytes + (row >> 3) * 256 + (kk >> 4) * 128 +
                       (row & 7) * 16 + (kk & 15);
    blob[dst] = static_cast<uint8_t>(blocks[b * 196 + i]);
  }
}

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// 16 pixels of one output channel: accumulators -> int8, optional residual add
// (golden_models.cpp:465-490), optional ReLU on the int8 value, zero for pixels >= n_valid.
template <bool RES, bool RES_FAST, bool SAT>
__device__ __forceinline__ uint4 ws_epi16(const WsParams& p, const uint32_t (&z)[16], const uint32_t (&u)[16], int bias, float sf,
                                          int relu_lo, int out_lo, const uint4& rbytes, int n_valid, uint32_t& sat) {
  uint32_t packed[4] = {0u, 0u, 0u, 0u};
  const uint32_t rw[4] = {rbytes.x, rbytes.y, rbytes.z, rbytes.w};
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int acc = max(static_cast<int>(z[e] + u[e]) + bias, relu_lo);
    const float f = __fmul_rn(__int2float_rn(acc), sf);
    int r8 = cvt_sat_s8(f);
    if constexpr (SAT) sat += (e < n_valid && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
    if constexpr (RES) {
      const int rv = static_cast<int>(static_cast<int8_t>((rw[e >> 2] >> (8 * (e & 3))) & 0xffu));
      const float a = __fmul_rn(__int2float_rn(r8), p.epi.res_scale_main);
      const float r = __fmul_rn(__int2float_rn(rv), p.epi.res_scale_res);
      const float s = __fadd_rn(a, r);
      float d;
      if constexpr (RES_FAST) {
        const float q0 = __fmul_rn(s, p.res_rcp);
        const float er = __fmaf_rn(-q0, p.epi.res_scale_out, s);
        d = __fmaf_rn(er, p.res_rcp, q0);
      } else {
        d = __fdiv_rn(s, p.epi.res_scale_out);
      }
      r8 = cvt_sat_s8(d);
    }
    const int q = e < n_valid ? max(r8, out_lo) : 0;
    packed[e >> 2] |= (static_cast<uint32_t>(q) & 0xffu) << (8 * (e & 3));
  }
  return make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

__global__ void __launch_bounds__(kWsThreads, 1) conv_ws_kernel(const __grid_constant__ WsLaunch L) {
  extern __shared__ uint8_t smem_dyn[];
  const WsParams& p = L.p;
  // operand areas need 1024-byte alignment (swizzle atoms): align the dynamic window by hand
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem = smem_dyn + (base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_full = bars;                          // [kWsMaxASlots] count 1 (+tx)
  uint64_t* a_empty = a_full + kWsMaxASlots;        // [kWsMaxASlots] count 1 (tcgen05.commit)
  uint64_t* w_full = a_empty + kWsMaxASlots;        // [kWsMaxChunks] count 1 (+tx)   (resident: one per chunk)
  uint64_t* w_empty = w_full + kWsMaxChunks;        // [kWsMaxWSlots] count 1
  uint64_t* acc_full = w_empty + kWsMaxWSlots;      // [2] count 1
  uint64_t* acc_empty = acc_full + 2;               // [2] count kWsEpiWarps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const uint32_t w_addr = base + kWsSmemBar;
  const uint32_t a_addr = w_addr + static_cast<uint32_t>(p.w_slots) * kWsChunkBytes;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t G = static_cast<uint32_t>(p.n_groups);
  const uint32_t g = blockIdx.x % G;                               // channel group of this CTA (weights stay put)
  const uint32_t tile0 = blockIdx.x / G, tile_step = gridDim.x / G;
  const uint32_t n_tiles = static_cast<uint32_t>(p.n_tiles);
  const uint32_t n_chunks = static_cast<uint32_t>(p.n_chunks);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWsMaxASlots; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kWsMaxChunks; ++s) mbar_init(&w_full[s], 1);
    for (int s = 0; s < kWsMaxWSlots; ++s) mbar_init(&w_empty[s], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kWsEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kWsWarpIssue) {
    tmem_alloc_dyn(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWsEpiWarps) {
    // =================================================================== epilogue: thread = output channel
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int co = static_cast<int>(g) * kWsCo + q * 32 + lane;
    const bool ch_ok = co < p.c_out;
    const float sf = ch_ok ? p.epi.chan_scale[co] : 0.f;
    const int bias = (ch_ok && p.epi.bias) ? p.epi.bias[co] : 0;
    const int relu_lo = (p.epi.flags & ACCEL_RELU) ? 0 : INT_MIN;
    const int out_lo = (p.epi.flags & ACCEL_RELU_OUT) ? 0 : -128;
    const bool sat_on = p.epi.sat_count != nullptr;
    const int resmode = !p.epi.residual ? 0 : (p.res_fast ? 1 : 2);
    const int n_c16 = p.N >> 4;                       // 16-pixel chunks per tile
    const int psh = p.P == 64 ? 6 : (p.P == 32 ? 5 : 4);
    const bool warp_has_ch = static_cast<int>(g) * kWsCo + q * 32 < p.c_out;
    uint32_t sat = 0, n = 0;
    for (uint32_t tt = tile0; tt < n_tiles; tt += tile_step, ++n) {
      const uint32_t img = fdiv(tt, p.d_tpi);
      const int y0 = static_cast<int>(tt - img * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
      const uint32_t ab = n & 1u;
      const uint32_t acc = tmem_base + lane_base + ab * kWsAccCols;
      const int64_t obase = static_cast<int64_t>(img) * p.image_stride + static_cast<int64_t>(co) * p.chan_stride;
      mbar_wait(&acc_full[ab], (n >> 1) & 1u);
      tc_fence_after();
      // chunk pairs alternate between the two warp sets: every thread stores 32 contiguous bytes
      for (int i = 0; i < n_c16 && warp_has_ch; ++i) {
        if (((i >> 1) & 1) != half) continue;
        const int p0 = i << 4;
        const int r = p0 >> psh, x0 = p0 & (p.P - 1);
        const int y = y0 + r;
        if (y >= p.H || x0 >= p.x_store_end) continue;     // warp-uniform
        uint32_t z[16], u[16];
        tmem_ld16(acc + p0, z);
        tmem_ld16(acc + (p.N - 2) + 1 + p0, u);
        const int64_t off = obase + static_cast<int64_t>(y) * p.out_pitch + x0;
        uint4 rb = make_uint4(0u, 0u, 0u, 0u);
        if (resmode && ch_ok) rb = ldg128(p.epi.residual + off);
        tmem_ld_wait();
        const int n_valid = min(16, p.W - x0);
        uint4 o;
        switch (resmode * 2 + (sat_on ? 1 : 0)) {
          case 0: o = ws_epi16<false, false, false>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
          case 1: o = ws_epi16<false, false, true>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
          case 2: o = ws_epi16<true, true, false>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
          case 3: o = ws_epi16<true, true, true>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
          case 4: o = ws_epi16<true, false, false>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
          default: o = ws_epi16<true, false, true>(p, z, u, bias, sf, relu_lo, out_lo, rb, n_valid, sat); break;
        }
        if (ch_ok) stg128(p.out + off, o);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
    if (sat_on) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, ch_ok ? sat : 0u);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
  } else if (warp == kWsWarpIssue) {
    // =================================================================== MMA issuer
    if (elect_one()) {
      const uint32_t idesc = idesc_i8_bmn(kWsCo, static_cast<uint32_t>(p.N));
      const uint64_t adesc0 = smem_desc_kmajor(0, 128, 256);
      const uint64_t bdesc0 = smem_desc_any(0, p.b_lbo, p.b_sbo, p.b_layout);
      const uint32_t a_hi = static_cast<uint32_t>(adesc0 >> 32), b_hi = static_cast<uint32_t>(bdesc0 >> 32);
      const uint32_t b_lo0 = static_cast<uint32_t>(bdesc0);
      const uint32_t a_lo0 = static_cast<uint32_t>(adesc0);
      const uint32_t row16 = p.row_stride >> 4;
      const uint32_t u_off = static_cast<uint32_t>(p.N - 2);
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, n = 0;
      for (uint32_t tt = tile0; tt < n_tiles; tt += tile_step, ++n) {
        const uint32_t ab = n & 1u;
        mbar_wait(&acc_empty[ab], ((n >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t z1 = tmem_base + ab * kWsAccCols;
        uint32_t z_on = 0u, u_on = 0u;                // first MMA into Z1 / U overwrites, the rest accumulate
        for (uint32_t j = 0; j < n_chunks; ++j) {
          uint32_t wslot;
          if (p.w_resident) {
            wslot = j;
            if (n == 0) mbar_wait(&w_full[j], 0u);
          } else {
            wslot = ws;
            mbar_wait(&w_full[ws], wph);
          }
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t mask = p.masks[g * kWsMaxChunks + j];
          const uint32_t wl = a_lo0 | (((w_addr + wslot * kWsChunkBytes) >> 4) & 0x3FFFu);
          const uint32_t xl = b_lo0 | (((a_addr + as * static_cast<uint32_t>(p.a_stage_bytes)) >> 4) & 0x3FFFu);
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (xl + kh * row16);
            // order kw = 1, 2, 0: the first U-type MMA of a tile is the unshifted one (covers U's first columns)
            if (mask & (1u << (kh * 3 + 1))) {
              mma_i8_ss(z1, (static_cast<uint64_t>(a_hi) << 32) | (wl + (kh * 3 + 1) * (kWsTapBytes >> 4)), bd, idesc, z_on);
              z_on = 1u;
            }
            if (mask & (1u << (kh * 3 + 2))) {
              mma_i8_ss(z1 + u_off, (static_cast<uint64_t>(a_hi) << 32) | (wl + (kh * 3 + 2) * (kWsTapBytes >> 4)), bd, idesc, u_on);
              u_on = 1u;
            }
            if (mask & (1u << (kh * 3 + 0))) {
              mma_i8_ss(z1 + u_off + 2, (static_cast<uint64_t>(a_hi) << 32) | (wl + (kh * 3 + 0) * (kWsTapBytes >> 4)), bd, idesc, u_on);
              u_on = 1u;
            }
          }
          mma_commit(&a_empty[as]);
          if (++as == static_cast<uint32_t>(p.a_slots)) { as = 0; aph ^= 1u; }
          if (!p.w_resident) {
            mma_commit(&w_empty[ws]);
            if (++ws == static_cast<uint32_t>(p.w_slots)) { ws = 0; wph ^= 1u; }
          }
        }
        mma_commit(&acc_full[ab]);
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // =================================================================== loader: weights (bulk) + activation tiles (TMA)
    if (elect_one()) {
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, n = 0;
      const uint8_t* wsrc = p.wblob + static_cast<size_t>(g) * n_chunks * kWsChunkBytes;
      for (uint32_t tt = tile0; tt < n_tiles; tt += tile_step, ++n) {
        const uint32_t img = fdiv(tt, p.d_tpi);
        const int y0 = static_cast<int>(tt - img * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
        for (uint32_t j = 0; j < n_chunks; ++j) {
          if (!p.w_resident || n == 0) {
            uint32_t wslot;
            if (p.w_resident) {
              wslot = j;
            } else {
              wslot = ws;
              mbar_wait(&w_empty[ws], wph ^ 1u);
              if (++ws == static_cast<uint32_t>(p.w_slots)) { ws = 0; wph ^= 1u; }
            }
            uint64_t* bar = p.w_resident ? &w_full[j] : &w_full[wslot];
            mbar_arrive_expect_tx(bar, kWsChunkBytes);
            const uint8_t* src = wsrc + static_cast<size_t>(j) * kWsChunkBytes;
            uint8_t* dst = smem + kWsSmemBar + wslot * kWsChunkBytes;
#pragma unroll
            for (int i = 0; i < 3; ++i) bulk_g2s(dst + i * (kWsChunkBytes / 3), src + i * (kWsChunkBytes / 3), kWsChunkBytes / 3, bar);
          }
          mbar_wait(&a_empty[as], aph ^ 1u);
          mbar_arrive_expect_tx(&a_full[as], static_cast<uint32_t>(p.a_box_bytes));
          tma_load_4d(a_addr + as * static_cast<uint32_t>(p.a_stage_bytes), &L.tmap, 0, static_cast<int>(j * kWsCk), y0 - 1,
                      static_cast<int>(img), &a_full[as]);
          if (++as == static_cast<uint32_t>(p.a_slots)) { as = 0; aph ^= 1u; }
        }
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWsWarpIssue) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
}

}  // namespace accel
