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

constexpr int kWsEpiWarps = 8;                     // warps 0-7: epilogue
constexpr int kWsWarpIssue = kWsEpiWarps;          // 8: MMA issuer (+ TMEM allocation)
constexpr int kWsWarpWeights = kWsWarpIssue + 1;   // 9: weight tiles by bulk copy (its own warp: it must never hold back
                                                   //    the activation loaders, or the two rings deadlock)
constexpr int kWsWarpLoad = kWsWarpWeights + 1;    // 10, 11: activation loaders (LDGSTS)
constexpr int kWsLoadWarps = 2;
constexpr int kWsLoadThreads = kWsLoadWarps * 32;  // 64
constexpr int kWsThreads = (kWsWarpLoad + kWsLoadWarps) * 32;   // 384
constexpr int kWsLoadOps = 8;                      // 16-byte copies per loader thread and stage (<= 512 per stage)
constexpr int kWsCo = 128;                         // output channels per group (TMEM lanes)
constexpr int kWsCk = 32;                          // input channels per chunk (one MMA K)
constexpr int kWsTapBytes = kWsCo * kWsCk;         // 4096
constexpr int kWsChunkBytes = 9 * kWsTapBytes;     // 36864
constexpr int kWsMaxChunks = 64;                   // Cin <= 2048
constexpr int kWsMaxGroups = 16;                   // Cout <= 2048
constexpr int kWsMaxChunks2 = 16, kWsMaxGroups2 = 8;      // the fused 1x1 / stride 2 branch (masks2): Cin <= 512, Cout <= 1024
constexpr int kWsRing9 = 4;                        // 3x3: weights resident when Cin <= 128, else a ring of this many 36 KB chunk slots
constexpr int kWsMaxWSlots = 32;                   // weight slots (1x1: 4 KB chunks - resident up to Cin 1024, else a ring)
constexpr int kWsMaxASlots = 16;
constexpr int kWsAccCols = 256;                    // per accumulator set: Z1 [0, N), U [N-2, 2N)
constexpr int kWsSmemBar = 2048;                   // barriers (132 x 8 bytes) + TMEM slot in front of the operand areas

struct WsParams {
  int32_t C, H, W, B;          // input geometry (output has the same H, W)
  int32_t P, R, N;             // pixels per staged row (16 / 32 / 64), output rows per tile, N = R * P
  int32_t n_chunks, n_groups, c_out;
  int32_t tiles_per_image, n_tiles;   // row tiles per image, B * tiles_per_image
  int32_t w_slots, w_resident, a_slots, a_stage_bytes, a_box_bytes;
  const int8_t* x;             // input tensor, rows of in_pitch bytes
  int32_t in_pitch;
  int32_t u_off, u_alias;      // U = Z1 + u_off.  N = 128 (or stride 2): u_off = N - 2, U's first two columns alias Z1's last two
                               // (padding pixels, exact zeros).  Smaller N: u_off = N, nothing aliases.
  int32_t twin;                // W <= 7 (16-pixel rows): a staged row holds row y of TWO images, [A0..A6 0 B0..B6 0]: the zero
                               // column between them is the padding of both, and a tile carries two images
  int32_t dual;                // c_out <= 64: items are pairs of pixel tiles, two interleaved M = 64 accumulators
  uint32_t b_layout, b_lbo, b_sbo, row_stride;
  FastDiv d_tpi;
  const uint8_t* wblob;        // [group][chunk][tap][4096]
  accel_epilogue epi;
  int32_t res_fast;            // residual: 0 IEEE divide, 1 exact 3-instruction sequence, 2 single multiply, 3 integer add (verified on the host)
  float res_rcp;
  int8_t* out;
  int32_t out_pitch;           // bytes between output rows
  int32_t chan_stride;         // H * out_pitch
  int32_t x_store_end;         // pixels of a row that are stored (W rounded up to 16)
  int64_t image_stride;        // c_out * chan_stride
  int32_t dbg;                 // developer aid (compiled in with -DACCEL_DEV=1 only): bit 0 = epilogue does no work, bit 1 = issuer issues no MMAs
  long long* timeline;         // developer aid: 16 clock64 stamps per CTA when non-null (accel_debug_set_timeline)
  uint16_t masks[kWsMaxGroups * kWsMaxChunks];
  // stride-2 mode (3x3 / stride 2 / pad 1, optionally fused with the 1x1 / stride 2 convolution that reads the same input:
  // the ResNet downsample).  Rows: the loader stages 2R+1 input rows and the B window of tap kh takes every second one.
  // Columns: computed at full resolution, the epilogue keeps the even ones.  The 1x1 convolution is one more tap on the
  // kh = 1 window, accumulated into V (columns [128, 128 + N), N <= 64) and written through its own epilogue.
  // Streamed weights (more than kWsMaxWSlots chunks) make every tile re-read the whole filter from L2, so there the tile
  // is as large as it can be (N <= 128, a whole 14 x 14 image): with the fused 1x1 the three accumulators then take
  // one 512-column set (acc_single: Z1/U in [0, 256), V at 256; the epilogue no longer overlaps the next tile's MMAs).
  int32_t stride, rows_in, Ho, Wo, w_chunk_bytes, has_ds, acc_single, v_col;
  const uint8_t* wblob2;       // [group][chunk][4096]
  accel_epilogue epi2;
  int8_t* out2;
  uint16_t masks2[kWsMaxGroups2 * kWsMaxChunks2];
  // pointwise mode (1x1 / stride 1 / pad 0: the bottleneck convolutions of ResNet-50): one tap, no halo rows (ypad = 0), Z1
  // only - the epilogue takes U as zero; the weight blob is [group][chunk][4096] (w_src_stride = 4096)
  int32_t pw, ypad, w_src_stride;
  // activation stages by TMA tensor tiles instead of LDGSTS (one elected thread; the tensor map of WsLaunch describes the NCHW input with
  // the dimensions ordered (x, c, y, n): box [P][32 channels][rows_in][1], swizzle atom = one image row of 8 channels)
  int32_t use_tma;
};
// Developer aids (stage-isolation flags, clock stamps) sit inside the loader / issuer / epilogue loops: they are compiled out
// unless the library is built with -DACCEL_DEV=1 (tools/ws_timeline.py, tools/ws_probe.py need that build).
#ifndef ACCEL_DEV
#define ACCEL_DEV 0
#endif
#ifndef ACCEL_WS_FENCE
#define ACCEL_WS_FENCE 0        // issuer-side fence.proxy.async between the cp.async stages and the MMAs: 0 none, 1 one per stage (see the issuer)
#endif
struct WsParams;
__device__ __forceinline__ int ws_dbg(const WsParams& p);
struct WsLaunch {
  alignas(64) CUtensorMap tmap;
  WsParams p;
};

__device__ __forceinline__ int ws_dbg(const WsParams& p) { return ACCEL_DEV ? p.dbg : 0; }
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
                                  int32_t taps, uint8_t* __restrict__ blob) {
  const int64_t b = blockIdx.x;
  if (b >= nnz) return;
  const int br = blk_row[b], bc = col_idx[b];
  for (int i = threadIdx.x; i < kBlock * kBlock; i += blockDim.x) {
    const int h = i / kBlock, w = i - h * kBlock;
    const int co = br * kBlock + h, k = bc * kBlock + w;
    if (co >= c_out || k >= c_in * taps) continue;
    const int c = k / taps, tap = k - c * taps;
    const int g = co / kWsCo, row = co - g * kWsCo, j = c / kWsCk, kk = c - j * kWsCk;
    const size_t dst = (static_cast<size_t>(g * n_chunks + j) * taps + tap) * kWsTapBytes + (row >> 3) * 256 + (kk >> 4) * 128 +
                       (row & 7) * 16 + (kk & 15);
    blob[dst] = static_cast<uint8_t>(blocks[b * 196 + i]);
  }
}

__device__ __forceinline__ uint32_t pin(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ int pin(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ uint64_t pin64(uint64_t v) { asm volatile("" : "+l"(v)); return v; }
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int B>
__device__ __forceinline__ float i2f_s8_byte(uint32_t w) {     // float(int8 at byte B of w): one I2F with a byte selector
  float f;
  if constexpr (B == 0) asm("cvt.rn.f32.s8 %0, %1;" : "=f"(f) : "r"(w));
  else asm("{\n\t.reg .b32 t;\n\tshr.u32 t, %1, %2;\n\tcvt.rn.f32.s8 %0, t;\n\t}" : "=f"(f) : "r"(w), "n"(8 * B));
  return f;
}

// Accumulator range [lo, hi] of one channel inside which the requant does not clip: f = RN(float(acc)) * sf with
// -128.5 <= f < 127.5 (the counting rule of the gather kernels' epilogue).  f is monotone in acc for sf > 0.
__device__ __forceinline__ void ws_sat_bounds(float sf, int& lo, int& hi) {
  if (!(sf > 0.f) || !(sf < 3.0e38f)) { lo = INT_MAX; hi = INT_MIN; return; }   // always take the exact count
  long long a = 0, b = INT_MAX;
  while (a < b) {
    const long long m = (a + b + 1) >> 1;
    if (__fmul_rn(__int2float_rn(static_cast<int>(m)), sf) < 127.5f) a = m; else b = m - 1;
  }
  hi = static_cast<int>(a);
  a = INT_MIN; b = 0;
  while (a < b) {
    const long long m = (a + b) >> 1;
    if (__fmul_rn(__int2float_rn(static_cast<int>(m)), sf) >= -128.5f) b = m; else a = m + 1;
  }
  lo = static_cast<int>(b);
}

// 16 pixels of one output channel: accumulators -> int8, optional residual add
// (golden_models.cpp:465-490), optional ReLU on the int8 value.  SAT: track the accumulator range of the chunk.
// cvt.rni.sat.s8.f32 leaves the int8 in the low byte of the register; the byte permutes below take that byte as it is
// (an int8_t return value would cost two more permutes per value to sign-extend).
__device__ __forceinline__ uint32_t cvt_sat_s8_raw(float f) {
  uint32_t v;
  asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(v) : "f"(f));
  return v;
}
__device__ __forceinline__ uint32_t pack4_b0(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {      // byte 0 of each
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
// relu_int8 on four packed int8 when relu_mask is all ones (0: leave them): clear the bytes whose sign bit is set
__device__ __forceinline__ uint32_t relu4_s8(uint32_t x, uint32_t relu_mask) {
  uint32_t sign;      // 0xFF in every byte whose sign bit is set (prmt's sign-replicate selectors; __byte_perm keeps only 3 selector bits)
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(sign) : "r"(x));
  return x & ~(sign & relu_mask);
}
struct WsEpiConst {        // per-thread (= per-channel) constants of the epilogue
  int bias, relu_lo, out_lo, lo_c, hi_c;
  float sf;
  // FAST path (see ws_epi16): bias + 0x4B400000, clamp bounds that implement relu_int32 + int8 saturation in the float
  // domain, and the thresholds beyond which a value counts as clipped
  int bias_m;
  float lo_f, hi_f, trip_lo, trip_hi;
};
// relu_int32 followed by the requant equals a clamp of the requantised value - on the side the sign of the factor picks
constexpr float kWsMagic = 12582912.0f;        // 1.5 * 2^23: float(kWsMagic + n) has n in its low mantissa bits for |n| < 2^22
__device__ __forceinline__ void ws_fast_consts(WsEpiConst& k, bool relu) {
  k.bias_m = k.bias + 0x4B400000;
  k.lo_f = -128.f; k.hi_f = 127.f; k.trip_lo = -128.5f; k.trip_hi = 127.5f;
  if (relu) {
    if (k.sf > 0.f) { k.lo_f = 0.f; k.trip_lo = -INFINITY; }            // acc < 0 -> 0: the negative side never clips
    else if (k.sf < 0.f) { k.hi_f = 0.f; k.trip_hi = INFINITY; }        // acc < 0 <=> product > 0 -> 0
    else { k.lo_f = k.hi_f = 0.f; k.trip_lo = -INFINITY; k.trip_hi = INFINITY; }
  }
}

// FAST (chosen on the host, accel_epilogue::acc_bound): no conversion instructions at all.  ncu on the layer1 kernels showed
// the XU pipe (I2F / F2I run there, 16 lanes per clock and SM) 75 % busy - the epilogue, not the tensor pipe, bounded them.
// With |accumulator + bias| < 2^22 guaranteed by the host (128 * the largest row L1 norm of the INT8 weights + the largest
// |bias|), float(acc) is the integer add  bits = z + u + (bias + 0x4B400000)  followed by an exact  - 1.5 * 2^23;  relu_int32
// and the int8 saturation become one clamp in the float domain (requant is monotone in the accumulator; the clamp bounds
// depend on the sign of the channel's factor), and round-to-nearest-even + float -> int is  + 1.5 * 2^23  again, whose low
// byte is the int8.  Same results bit for bit (the multiply by sf is the same single float32 multiply).
template <int RESMODE, bool SAT, bool FAST>
__device__ __forceinline__ uint4 ws_epi16(const WsParams& p, const uint32_t (&z)[16], const uint32_t (&u)[16], const WsEpiConst& k,
                                          const uint4& rbytes, int& amin, int& amax, float& fmn, float& fmx) {
  uint32_t packed[4];
  const uint32_t rw[4] = {rbytes.x, rbytes.y, rbytes.z, rbytes.w};
  const uint32_t relu_mask = k.out_lo == 0 ? 0xFFFFFFFFu : 0u;        // out_lo is 0 (relu_int8) or -128 (none)
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint32_t q[4];
    if constexpr (FAST) {
      static_assert(RESMODE == 0 || RESMODE == 4, "the conversion-free epilogue covers no residual / the integer residual add");
      float f[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int e = 4 * w + b;
        const float t = __int_as_float(static_cast<int>(z[e] + u[e]) + k.bias_m);
        f[b] = __fmul_rn(__fadd_rn(t, -kWsMagic), k.sf);
      }
      if constexpr (SAT) {
        fmx = fmaxf(fmaxf(fmx, f[0]), fmaxf(f[1], fmaxf(f[2], f[3])));
        fmn = fminf(fminf(fmn, f[0]), fminf(f[1], fminf(f[2], f[3])));
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) q[b] = __float_as_uint(__fadd_rn(fminf(fmaxf(f[b], k.lo_f), k.hi_f), kWsMagic));
    } else {
      int acc4[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) acc4[b] = max(static_cast<int>(z[4 * w + b] + u[4 * w + b]) + k.bias, k.relu_lo);
      if constexpr (SAT) {                   // range of the chunk, two values per three-input min / max
        amax = __vimax3_s32(__vimax3_s32(amax, acc4[0], acc4[1]), acc4[2], acc4[3]);
        amin = __vimin3_s32(__vimin3_s32(amin, acc4[0], acc4[1]), acc4[2], acc4[3]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int acc = acc4[b];
        const float f = __fmul_rn(__int2float_rn(acc), k.sf);
        q[b] = cvt_sat_s8_raw(f);
        if constexpr (RESMODE != 0 && RESMODE != 4) {
          const float rf = b == 0 ? i2f_s8_byte<0>(rw[w]) : b == 1 ? i2f_s8_byte<1>(rw[w]) : b == 2 ? i2f_s8_byte<2>(rw[w])
                                                                                                       : i2f_s8_byte<3>(rw[w]);
          const float a = __fmul_rn(i2f_s8_byte<0>(q[b]), p.epi.res_scale_main);
          const float r = __fmul_rn(rf, p.epi.res_scale_res);
          const float sm = __fadd_rn(a, r);
          float d;
          if constexpr (RESMODE == 3) {      // the multiply alone already rounds to the reference's int8 for every pair
            d = __fmul_rn(sm, p.res_rcp);
          } else if constexpr (RESMODE == 1) {      // exact for every (int8, int8) pair: verified on the host
            const float q0 = __fmul_rn(sm, p.res_rcp);
            const float er = __fmaf_rn(-q0, p.epi.res_scale_out, sm);
            d = __fmaf_rn(er, p.res_rcp, q0);
          } else {
            d = __fdiv_rn(sm, p.epi.res_scale_out);
          }
          q[b] = cvt_sat_s8_raw(d);
        }
      }
    }
    uint32_t pk = pack4_b0(q[0], q[1], q[2], q[3]);
    // matched scales: the host checked that the reference's float sequence equals the saturating integer sum for all
    // 65 536 (main, residual) pairs - four of them per packed saturating add
    if constexpr (RESMODE == 4) pk = __vaddss4(pk, rw[w]);
    packed[w] = relu4_s8(pk, relu_mask);
  }
  return make_uint4(packed[0], packed[1], packed[2], packed[3]);
}
// exact count of clipped values among the first n_valid pixels of a chunk (rare path)
__device__ __forceinline__ uint32_t ws_sat_recount(const uint32_t (&z)[16], const uint32_t (&u)[16], int bias, float sf, int relu_lo,
                                                   int n_valid) {
  uint32_t c = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int acc = max(static_cast<int>(z[e] + u[e]) + bias, relu_lo);
    const float f = __fmul_rn(__int2float_rn(acc), sf);
    c += (e < n_valid && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
  }
  return c;
}

struct WsChunk {           // one 16-pixel chunk of this thread's channel
  int p0;                  // first pixel (TMEM column) of the chunk inside the tile
  int n_valid;             // pixels of the chunk inside the image row
  bool live;               // warp-uniform: somebody stores it
  bool lane_ok;            // this thread stores it
  int64_t off;             // element offset of the chunk in the output / residual tensor
};
template <bool PW>
__device__ __forceinline__ void ws_chunk_load(const WsParams& p, uint32_t acc, const WsChunk& c, uint32_t (&z)[16], uint32_t (&u)[16]) {
  tmem_ld16(acc + c.p0, z);
  if constexpr (PW) {
#pragma unroll
    for (int e = 0; e < 16; ++e) u[e] = 0u;
  } else {
    tmem_ld16(acc + p.u_off + 1 + c.p0, u);
  }
}
template <int RESMODE, bool SAT, bool FAST>
__device__ __forceinline__ void ws_chunk_finish(const WsParams& p, const WsChunk& c, const uint32_t (&z)[16], const uint32_t (&u)[16],
                                                const uint4& rb, const WsEpiConst& k, uint32_t& sat) {
  int amin = INT_MAX, amax = INT_MIN;
  float fmn = 0.f, fmx = 0.f;
  uint4 o = ws_epi16<RESMODE, SAT, FAST>(p, z, u, k, rb, amin, amax, fmn, fmx);
  if constexpr (SAT) {
    const bool trip = FAST ? (fmx >= k.trip_hi || fmn < k.trip_lo) : (amax > k.hi_c || amin < k.lo_c);
    if (c.lane_ok && trip) sat += ws_sat_recount(z, u, k.bias, k.sf, k.relu_lo, c.n_valid);
  }
  if (c.n_valid < 16) {      // pixels beyond the image width are stored as zeros (the row padding stays zero)
    uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int nv = c.n_valid - 4 * w;
      ow[w] = nv >= 4 ? ow[w] : (nv <= 0 ? 0u : (ow[w] & ((1u << (8 * nv)) - 1u)));
    }
    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  if (c.lane_ok) stg128(p.out + c.off, o);
}

// this warp's k-th chunk of a tile: chunk pairs alternate between the two warp sets (32 contiguous bytes per thread)
struct WsEpiGeom {
  int n_c16, psh, half, y0;
  bool ch_ok, t_ok;
  int64_t obase;
};
__device__ __forceinline__ bool ws_chunk_at(const WsParams& p, const WsEpiGeom& g, int k, WsChunk& c) {
  const int i = 2 * (g.half + 2 * (k >> 1)) + (k & 1);
  if (i >= g.n_c16) return false;
  c.p0 = i << 4;
  const int r = c.p0 >> g.psh, x0 = c.p0 & (p.P - 1);
  c.lane_ok = g.ch_ok && g.t_ok && (g.y0 + r) < p.H && x0 < p.x_store_end;
  c.live = __any_sync(0xffffffffu, c.lane_ok);
  c.off = g.obase + static_cast<int64_t>(g.y0 + r) * p.out_pitch + x0;
  c.n_valid = max(0, min(16, p.W - x0));
  return true;
}

// All chunks of one tile for this warp (at most kWsMaxMyChunks).  The residual bytes were requested before the tile's
// accumulators were complete (ws_epi_prefetch); the TMEM loads of chunk k + 1 are in flight while chunk k is
// converted and stored.
constexpr int kWsMaxMyChunks = 4;          // N <= 128: 8 chunks of 16 pixels, split between the two warp sets
template <int RESMODE>
__device__ __forceinline__ void ws_epi_prefetch(const WsParams& p, const WsEpiGeom& g, uint4 (&rpre)[kWsMaxMyChunks]) {
#pragma unroll
  for (int k = 0; k < kWsMaxMyChunks; ++k) {
    rpre[k] = make_uint4(0u, 0u, 0u, 0u);
    if constexpr (RESMODE != 0) {
      WsChunk c;
      if (ws_chunk_at(p, g, k, c) && c.lane_ok) rpre[k] = ldg128(p.epi.residual + c.off);
    }
  }
}
template <int RESMODE, bool SAT, bool FAST, bool PW>
__device__ __forceinline__ void ws_epi_tile(const WsParams& p, uint32_t acc, const WsEpiGeom& g, const WsEpiConst& kc,
                                            const uint4 (&rpre)[kWsMaxMyChunks], uint32_t& sat) {
  uint32_t za[16], ua[16], zb[16], ub[16];
  WsChunk ca, cb;
  bool has_a = ws_chunk_at(p, g, 0, ca), has_b = false;
  if (has_a) ws_chunk_load<PW>(p, acc, ca, za, ua);
#pragma unroll
  for (int k = 0; k < kWsMaxMyChunks; k += 2) {
    if (!has_a) break;
    tmem_ld_wait();
    has_b = ws_chunk_at(p, g, k + 1, cb);
    if (has_b) ws_chunk_load<PW>(p, acc, cb, zb, ub);
    if (ca.live) ws_chunk_finish<RESMODE, SAT, FAST>(p, ca, za, ua, rpre[k], kc, sat);
    if (!has_b) break;
    tmem_ld_wait();
    has_a = k + 2 < kWsMaxMyChunks && ws_chunk_at(p, g, k + 2, ca);
    if (has_a) ws_chunk_load<PW>(p, acc, ca, za, ua);
    if (cb.live) ws_chunk_finish<RESMODE, SAT, FAST>(p, cb, zb, ub, rpre[k + 1], kc, sat);
  }
  tmem_ld_wait();
}

__device__ __forceinline__ void stg64(void* p, const uint2& v) {
  asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// Twin tiles (WsParams::twin): every 16-pixel chunk is one row of two images, 8 bytes each.  Unpipelined: these layers
// are 7x7 images, the epilogue is a small part of them.
__device__ __forceinline__ uint2 ldg64(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
// ---- stride-2 epilogue: 16 full-resolution pixels -> the 8 even ones -> 8 output bytes
template <bool SAT, bool FAST>
__device__ __forceinline__ uint2 ws_epi8_even(const uint32_t (&z)[16], const uint32_t* u, const WsEpiConst& k, int n_valid, bool lane_ok,
                                              uint32_t& sat) {
  uint32_t q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if constexpr (FAST) {        // conversion-free arithmetic, see ws_epi16
      const float t = __int_as_float(static_cast<int>(z[2 * e] + (u ? u[2 * e] : 0u)) + k.bias_m);
      const float f = __fmul_rn(__fadd_rn(t, -kWsMagic), k.sf);
      if constexpr (SAT) sat += (lane_ok && e < n_valid && (f >= k.trip_hi || f < k.trip_lo)) ? 1u : 0u;
      q[e] = __float_as_uint(__fadd_rn(fminf(fmaxf(f, k.lo_f), k.hi_f), kWsMagic));
    } else {
      const int acc = max(static_cast<int>(z[2 * e] + (u ? u[2 * e] : 0u)) + k.bias, k.relu_lo);
      const float f = __fmul_rn(__int2float_rn(acc), k.sf);
      if constexpr (SAT) sat += (lane_ok && e < n_valid && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
      q[e] = cvt_sat_s8_raw(f);
    }
  }
  // bytes past the row end are stored as zeros (the row padding stays zero)
  const uint32_t relu_mask = k.out_lo == 0 ? 0xFFFFFFFFu : 0u;
  const uint32_t keep_lo = n_valid >= 4 ? 0xFFFFFFFFu : (n_valid <= 0 ? 0u : (1u << (8 * n_valid)) - 1u);
  const uint32_t keep_hi = n_valid >= 8 ? 0xFFFFFFFFu : (n_valid <= 4 ? 0u : (1u << (8 * (n_valid - 4))) - 1u);
  return make_uint2(relu4_s8(pack4_b0(q[0], q[1], q[2], q[3]), relu_mask) & keep_lo,
                    relu4_s8(pack4_b0(q[4], q[5], q[6], q[7]), relu_mask) & keep_hi);
}

// The epilogue role of one warp for the whole launch (instantiated per variant: the variant is chosen once, outside the loop).
struct WsEpiRole {
  uint32_t item0, item_step, n_items, n_tiles, dual, sub, tmem_acc;   // tmem_acc: TMEM address of accumulator set 0, this warp's lanes
  int co, half, n_c16, psh;
  bool ch_ok, warp_has_ch;
  uint64_t* acc_full;
  uint64_t* acc_empty;
};
template <int RESMODE, bool SAT, bool FAST, bool PW>
__device__ __forceinline__ uint32_t ws_epi_loop(const WsParams& p, const WsEpiRole& r, const WsEpiConst& kc, int lane) {
  uint32_t sat = 0, n = 0;
  for (uint32_t it = r.item0; it < r.n_items; it += r.item_step, ++n) {
    const uint32_t tt = r.dual ? 2u * it + r.sub : it;
    const bool t_ok = tt < r.n_tiles;
    const uint32_t tc = t_ok ? tt : r.n_tiles - 1u;
    const uint32_t img = fdiv(tc, p.d_tpi);
    const uint32_t ab = n & 1u;
    WsEpiGeom eg;
    eg.n_c16 = r.n_c16; eg.psh = r.psh; eg.half = r.half; eg.ch_ok = r.ch_ok; eg.t_ok = t_ok;
    eg.y0 = static_cast<int>(tc - img * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
    eg.obase = static_cast<int64_t>(img) * p.image_stride + static_cast<int64_t>(r.co) * p.chan_stride;
    uint4 rpre[kWsMaxMyChunks];
    if (r.warp_has_ch && !(ws_dbg(p) & 1)) ws_epi_prefetch<RESMODE>(p, eg, rpre);      // residual bytes: before the MMAs are done
    mbar_wait(&r.acc_full[ab], (n >> 1) & 1u);
    tc_fence_after();
    if (r.warp_has_ch && !(ws_dbg(p) & 1)) ws_epi_tile<RESMODE, SAT, FAST, PW>(p, r.tmem_acc + ab * kWsAccCols, eg, kc, rpre, sat);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.acc_empty[ab]);
  }
  return sat;
}

template <int RESMODE, bool SAT, bool FAST, bool PW>
__device__ __forceinline__ uint32_t ws_epi_loop_twin(const WsParams& p, const WsEpiRole& r, const WsEpiConst& kc, int lane) {
  uint32_t sat = 0, n = 0;
  const uint64_t keep64 = p.W >= 8 ? ~0ull : ((1ull << (8 * p.W)) - 1ull);
  const uint32_t keep_lo = static_cast<uint32_t>(keep64), keep_hi = static_cast<uint32_t>(keep64 >> 32);
  for (uint32_t it = r.item0; it < r.n_items; it += r.item_step, ++n) {
    const uint32_t pr = fdiv(it, p.d_tpi);
    const int y0 = static_cast<int>(it - pr * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
    const uint32_t img_a = 2u * pr;
    const bool ok_a = r.ch_ok, ok_b = r.ch_ok && img_a + 1u < static_cast<uint32_t>(p.B);
    const uint32_t ab = n & 1u;
    const uint32_t acc = r.tmem_acc + ab * kWsAccCols;
    const int64_t obase = static_cast<int64_t>(img_a) * p.image_stride + static_cast<int64_t>(r.co) * p.chan_stride;
    mbar_wait(&r.acc_full[ab], (n >> 1) & 1u);
    tc_fence_after();
    if (r.warp_has_ch && !(ws_dbg(p) & 1)) {
      for (int k = 0; k < kWsMaxMyChunks; ++k) {
        const int i = 2 * (r.half + 2 * (k >> 1)) + (k & 1);         // chunk = staged row i of the tile
        if (i >= r.n_c16) break;
        if (y0 + i >= p.H) continue;
        const int64_t off_a = obase + static_cast<int64_t>(y0 + i) * p.out_pitch, off_b = off_a + p.image_stride;
        uint32_t z[16], u[16];
        tmem_ld16(acc + 16 * i, z);
        if constexpr (PW) {
#pragma unroll
          for (int e = 0; e < 16; ++e) u[e] = 0u;
        } else {
          tmem_ld16(acc + p.u_off + 1 + 16 * i, u);
        }
        uint4 rb = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (RESMODE != 0) {
          if (ok_a) { const uint2 t = ldg64(p.epi.residual + off_a); rb.x = t.x; rb.y = t.y; }
          if (ok_b) { const uint2 t = ldg64(p.epi.residual + off_b); rb.z = t.x; rb.w = t.y; }
        }
        tmem_ld_wait();
        int amin = INT_MAX, amax = INT_MIN;
        float fmn = 0.f, fmx = 0.f;
        const uint4 o = ws_epi16<RESMODE, SAT, FAST>(p, z, u, kc, rb, amin, amax, fmn, fmx);
        if constexpr (SAT) {
          const bool trip = FAST ? (fmx >= kc.trip_hi || fmn < kc.trip_lo) : (amax > kc.hi_c || amin < kc.lo_c);
          if (r.ch_ok && trip) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int a = max(static_cast<int>(z[e] + u[e]) + kc.bias, kc.relu_lo);
              const float f = __fmul_rn(__int2float_rn(a), kc.sf);
              sat += ((e & 7) < p.W && (e < 8 ? ok_a : ok_b) && !(f < 127.5f && f >= -128.5f)) ? 1u : 0u;
            }
          }
        }
        if (ok_a) stg64(p.out + off_a, make_uint2(o.x & keep_lo, o.y & keep_hi));
        if (ok_b) stg64(p.out + off_b, make_uint2(o.z & keep_lo, o.w & keep_hi));
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.acc_empty[ab]);
  }
  return sat;
}

template <bool SAT, bool FAST>
__device__ __forceinline__ uint32_t ws_epi_loop_s2(const WsParams& p, const WsEpiRole& r, const WsEpiConst& kc, const WsEpiConst& kc2,
                                                   int lane) {
  uint32_t sat = 0, n = 0;
  for (uint32_t it = r.item0; it < r.n_items; it += r.item_step, ++n) {
    const uint32_t img = fdiv(it, p.d_tpi);
    const int y0 = static_cast<int>(it - img * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
    const uint32_t ab = p.acc_single ? 0u : (n & 1u);
    const uint32_t acc = r.tmem_acc + ab * kWsAccCols;
    const int64_t obase = static_cast<int64_t>(img) * p.image_stride + static_cast<int64_t>(r.co) * p.chan_stride;
    mbar_wait(&r.acc_full[ab], p.acc_single ? (n & 1u) : ((n >> 1) & 1u));
    tc_fence_after();
    if (r.warp_has_ch && !(ws_dbg(p) & 1)) {
      for (int k = 0; k < kWsMaxMyChunks; ++k) {
        const int i = 2 * (r.half + 2 * (k >> 1)) + (k & 1);
        if (i >= r.n_c16) break;
        const int p0 = i << 4;
        const int row = p0 >> r.psh, x0 = (p0 & (p.P - 1)) >> 1;              // output column of the chunk's first pixel
        if (y0 + row >= p.Ho || x0 >= p.x_store_end) continue;                // warp-uniform
        uint32_t z[16], u[16], v[16];
        tmem_ld16(acc + p0, z);
        tmem_ld16(acc + p.u_off + 1 + p0, u);
        if (p.has_ds) tmem_ld16(acc + static_cast<uint32_t>(p.v_col) + p0, v);
        tmem_ld_wait();
        const int n_valid = max(0, min(8, p.Wo - x0));
        const int64_t off = obase + static_cast<int64_t>(y0 + row) * p.out_pitch + x0;
        const uint2 o = ws_epi8_even<SAT, FAST>(z, u, kc, n_valid, r.ch_ok, sat);
        if (r.ch_ok) stg64(p.out + off, o);
        if (p.has_ds) {
          const uint2 o2 = ws_epi8_even<SAT, FAST>(v, nullptr, kc2, n_valid, r.ch_ok, sat);
          if (r.ch_ok) stg64(p.out2 + off, o2);
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.acc_empty[ab]);
  }
  return sat;
}

// One chunk (32 input channels) of MMAs: nine taps (+ the fused 1x1 / stride 2 tap) from one staged activation tile.
// FIRST: the first chunk of a tile - the host forces all nine taps on, and the first MMA into Z1 / U / V overwrites.  Later
// chunks issue the taps of `mask` and always accumulate.  Tap order per kh: kw = 1 into Z1, then the two U taps - with
// aliasing kw = 0 (U + 2) first (it overwrites; the aliased columns keep Z1's zeros), without (twin tiles) kw = 2 (U + 0).
struct WsIssue {
  bool leader;
  uint32_t a_hi, b_hi, idesc, row16, u_off, v_col;
};
template <bool FIRST, bool ALIAS>
__device__ __forceinline__ void ws_issue_chunk(const WsIssue& I, uint32_t z1, uint32_t wl, uint32_t xl, uint32_t mask, bool ds) {
  constexpr uint32_t kTap16 = kWsTapBytes >> 4;
  const uint64_t ah = static_cast<uint64_t>(I.a_hi) << 32, bh = static_cast<uint64_t>(I.b_hi) << 32;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const uint64_t bd = bh | (xl + kh * I.row16);
    const uint32_t first = (FIRST && kh == 0) ? 0u : 1u;          // accumulate flag of the first Z1 / U MMA of a tile
    if (I.leader && (FIRST || (mask & (1u << (kh * 3 + 1))))) mma_i8_ss(z1, ah | (wl + (kh * 3 + 1) * kTap16), bd, I.idesc, first);
    constexpr int kwa = ALIAS ? 0 : 2, kwb = ALIAS ? 2 : 0;       // kwa overwrites on the first chunk
    const uint32_t ua = z1 + I.u_off + (ALIAS ? 2u : 0u), ub = z1 + I.u_off + (ALIAS ? 0u : 2u);
    if (I.leader && (FIRST || (mask & (1u << (kh * 3 + kwa))))) mma_i8_ss(ua, ah | (wl + (kh * 3 + kwa) * kTap16), bd, I.idesc, first);
    if (I.leader && (FIRST || (mask & (1u << (kh * 3 + kwb))))) mma_i8_ss(ub, ah | (wl + (kh * 3 + kwb) * kTap16), bd, I.idesc, 1u);
    if (kh == 1 && ds && I.leader) mma_i8_ss(z1 + I.v_col, ah | (wl + 9 * kTap16), bd, I.idesc, FIRST ? 0u : 1u);
  }
}

// One instantiation per (tile mode, residual mode, saturation counting): every launch runs exactly one epilogue variant, so
// the others cost it neither registers nor instruction-cache space (measured: the twin paths inside one big kernel slowed
// the layer1 convolutions from 63 to 84 us).  MODE 0 = stride 1, 1 = stride 2 (+ fused 1x1), 2 = twin tiles.
// MODE 3 / 4 = the pointwise (1x1 / stride 1 / pad 0) forms of 0 / 2: compile-time as well - a run-time flag inside the
// issuer and epilogue loops of the 3x3 kernels costs them several per cent (same-box A/B).
constexpr int kWsModeS1 = 0, kWsModeS2 = 1, kWsModeTwin = 2, kWsModePw = 3, kWsModeTwinPw = 4;
template <int MODE, int RESMODE, bool SAT, bool FAST>
__global__ void __launch_bounds__(kWsThreads, 1) conv_ws_kernel(const __grid_constant__ WsLaunch L) {
  constexpr bool TWIN = MODE == kWsModeTwin || MODE == kWsModeTwinPw;
  constexpr bool PW = MODE == kWsModePw || MODE == kWsModeTwinPw;
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
  const uint32_t a_addr = w_addr + static_cast<uint32_t>(p.w_slots) * static_cast<uint32_t>(p.w_chunk_bytes);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t G = static_cast<uint32_t>(p.n_groups);
  const uint32_t g = blockIdx.x % G;                               // channel group of this CTA (weights stay put)
  const uint32_t item0 = blockIdx.x / G, item_step = gridDim.x / G;
  const uint32_t n_tiles = static_cast<uint32_t>(p.n_tiles);
  const uint32_t n_chunks = static_cast<uint32_t>(p.n_chunks);
  // dual: <= 64 output channels.  An item is a PAIR of pixel tiles; each is an M = 64 accumulator that occupies
  // 16 lanes of every TMEM lane quadrant (tile s at lanes 16 s .. 16 s + 15): all 128 lanes carry results.
  const uint32_t dual = static_cast<uint32_t>(p.dual);
  const uint32_t n_items = dual ? (n_tiles + 1u) >> 1 : n_tiles;

  // The next kernel on the stream may take over SMs as CTAs of this grid retire (its prologue and resident weight
  // load then overlap our tail); everything here that reads what the previous kernel wrote sits behind griddep_wait().
  long long* const tl = (ACCEL_DEV && p.timeline) ? p.timeline + static_cast<size_t>(blockIdx.x) * 128 : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = clock64(); long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); tl[14] = g; }   // 0: CTA entry (14: ns)
  griddep_launch();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWsMaxASlots; ++s) { mbar_init(&a_full[s], p.use_tma ? 1 : kWsLoadThreads); mbar_init(&a_empty[s], 1); }
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
  if (tl && threadIdx.x == 0) tl[1] = clock64();                       // 1: prologue done (barriers, TMEM)

  if (warp < kWsEpiWarps) {
    // =================================================================== epilogue: thread = output channel
    griddep_wait();
    if (tl && threadIdx.x == 0) tl[2] = clock64();                     // 2: epilogue past griddepcontrol.wait          // residuals, the output buffer (still being read) and the counters belong to earlier kernels
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sub = dual ? static_cast<uint32_t>(lane >> 4) : 0u;          // which tile of the pair
    const int co = dual ? q * 16 + (lane & 15) : static_cast<int>(g) * kWsCo + q * 32 + lane;
    const bool ch_ok = co < p.c_out;
    const float sf = ch_ok ? p.epi.chan_scale[co] : 0.f;
    const int bias = (ch_ok && p.epi.bias) ? p.epi.bias[co] : 0;
    constexpr bool sat_on = SAT;
    WsEpiConst kc;
    kc.bias = bias; kc.sf = sf;
    kc.relu_lo = (p.epi.flags & ACCEL_RELU) ? 0 : INT_MIN;
    kc.out_lo = (p.epi.flags & ACCEL_RELU_OUT) ? 0 : -128;
    kc.lo_c = INT_MIN; kc.hi_c = INT_MAX;
    if (sat_on && ch_ok && !FAST) ws_sat_bounds(sf, kc.lo_c, kc.hi_c);
    ws_fast_consts(kc, (p.epi.flags & ACCEL_RELU) != 0);
    WsEpiRole er;
    er.item0 = item0; er.item_step = item_step; er.n_items = n_items; er.n_tiles = n_tiles; er.dual = dual; er.sub = sub;
    er.tmem_acc = tmem_base + lane_base;
    er.co = co; er.half = half; er.n_c16 = p.N >> 4; er.psh = p.P == 64 ? 6 : (p.P == 32 ? 5 : 4);
    er.ch_ok = ch_ok;
    er.warp_has_ch = dual ? (q * 16 < p.c_out) : (static_cast<int>(g) * kWsCo + q * 32 < p.c_out);
    er.acc_full = acc_full; er.acc_empty = acc_empty;
    uint32_t sat = 0;
    if constexpr (MODE == kWsModeS2) {
      WsEpiConst kc2 = kc;
      if (p.has_ds) {
        kc2.sf = ch_ok ? p.epi2.chan_scale[co] : 0.f;
        kc2.bias = (ch_ok && p.epi2.bias) ? p.epi2.bias[co] : 0;
        kc2.relu_lo = (p.epi2.flags & ACCEL_RELU) ? 0 : INT_MIN;
        kc2.out_lo = (p.epi2.flags & ACCEL_RELU_OUT) ? 0 : -128;
        ws_fast_consts(kc2, (p.epi2.flags & ACCEL_RELU) != 0);
      }
      sat = ws_epi_loop_s2<SAT, FAST>(p, er, kc, kc2, lane);
    } else if constexpr (TWIN) {
      sat = ws_epi_loop_twin<RESMODE, SAT, FAST, PW>(p, er, kc, lane);
    } else {
      sat = ws_epi_loop<RESMODE, SAT, FAST, PW>(p, er, kc, lane);
    }
    if (tl && threadIdx.x == 0) tl[3] = clock64();                     // 3: epilogue loop done (warp 0)
    if (sat_on) {
      const uint32_t wsum = __reduce_add_sync(0xffffffffu, sat);
      if (lane == 0 && wsum) atomicAdd(p.epi.sat_count, static_cast<unsigned long long>(wsum));
    }
  } else if (warp == kWsWarpIssue) {
    // =================================================================== MMA issuer
    // One elected thread.  Measured (tools/ws_timeline.py, stage isolation): the first version of this loop kept the
    // overwrite / accumulate state of Z1, U and V in variables and tested the tap masks in every chunk; every descriptor,
    // accumulator address and predicate then travelled through R2UR and the ~170 instructions per chunk of nine 64-cycle
    // MMAs took ~660 cycles - more than the 576 cycles the tensor pipe needs, i.e. every convolution was bound by this one
    // thread.  Now the first chunk of a tile (all taps, fixed overwrite pattern) and the later chunks (tap masks, always
    // accumulate) are separate code and the descriptors are base + slot * stride adds, which ptxas keeps in the uniform
    // datapath (UTCIMMA reads uniform registers).
    if (elect_one()) {
      const bool leader = true;
      const uint32_t idesc = idesc_i8_bmn(dual ? 64u : static_cast<uint32_t>(kWsCo), static_cast<uint32_t>(p.N));
      const uint64_t adesc0 = smem_desc_kmajor(0, 128, 256);
      const uint64_t bdesc0 = smem_desc_any(0, p.b_lbo, p.b_sbo, p.b_layout);
      WsIssue I;
      I.leader = leader;
      I.a_hi = static_cast<uint32_t>(adesc0 >> 32); I.b_hi = static_cast<uint32_t>(bdesc0 >> 32);
      I.idesc = idesc; I.row16 = p.row_stride >> 4; I.u_off = static_cast<uint32_t>(p.u_off); I.v_col = static_cast<uint32_t>(p.v_col);
      // shared memory ends below 256 KB: (address >> 4) fits the 14-bit field, so slot addressing is a plain add
      const uint32_t wl0 = static_cast<uint32_t>(adesc0) + (w_addr >> 4), w_step = static_cast<uint32_t>(p.w_chunk_bytes) >> 4;
      const uint32_t xl0 = static_cast<uint32_t>(bdesc0) + (a_addr >> 4), a_step = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
      const uint32_t a_slots = static_cast<uint32_t>(p.a_slots), w_slots = static_cast<uint32_t>(p.w_slots);
      const bool w_resident = p.w_resident != 0, has_ds = p.has_ds != 0, no_mma = (ws_dbg(p) & 2) != 0, fence = TWIN || ACCEL_WS_FENCE == 1 || (ACCEL_DEV && (p.dbg & 8) != 0);
      const bool one_set = MODE == kWsModeS2 && p.acc_single;
      constexpr bool alias = !TWIN;
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, n = 0, st_i = 0;
      for (uint32_t it = item0; it < n_items; it += item_step, ++n) {
        const uint32_t ab = one_set ? 0u : (n & 1u);
        mbar_wait(&acc_empty[ab], (one_set ? (n & 1u) : ((n >> 1) & 1u)) ^ 1u);
        tc_fence_after();
        const uint32_t n_sub = (dual && 2u * it + 1u < n_tiles) ? 2u : 1u;
        for (uint32_t sub = 0; sub < n_sub; ++sub) {
          const uint32_t z1 = tmem_base + ab * kWsAccCols + ((sub * 16u) << 16);
          for (uint32_t j = 0; j < n_chunks; ++j) {
            uint32_t wslot;
            if (w_resident) {
              wslot = j;
              if (n == 0 && sub == 0) mbar_wait(&w_full[j], 0u);
            } else {
              wslot = ws;
              mbar_wait(&w_full[ws], wph);
            }
            if (tl && leader && n == 0 && sub == 0 && j == 0) tl[4] = clock64();   // 4: issuer has the first weights
            // Proxy fence.  Stages that arrive by TMA are written and read in the async proxy: nothing to order.  The twin
            // tiles are still written by cp.async (generic proxy) and read by tcgen05.mma (async proxy); read strictly, the PTX memory
            // model asks for a fence.proxy.async between the mbarrier wait that makes those writes visible and the MMAs, so the twin
            // kernels issue one per stage here (compile-time: TWIN).  What it costs was measured when every layer still used cp.async
            // (whole ResNet-18, batch 256, one box): no fence 262.1 k img/s, one per stage 255.2 k (-2.6 %), one per group of up to four
            // arrived stages (mbarrier.test_wait look-ahead) 245.2 k; a run-time switch inside this loop costs 6 % by itself.  The
            // developer path that loads non-twin stages with cp.async (ACCEL_WS_TMA=0) is fenced only in -DACCEL_WS_FENCE=1 builds.
            mbar_wait(&a_full[as], aph);
            if (fence) fence_proxy_async_smem();
            if (tl && st_i >= 32 && st_i < 64) tl[64 + (st_i - 32)] = clock64();      // issuer saw stage st_i
            ++st_i;
            if (tl && leader && n == 0 && sub == 0 && j == 0) tl[5] = clock64();   // 5: issuer has the first activation stage
            tc_fence_after();
            const uint32_t wl = wl0 + wslot * w_step, xl = xl0 + as * a_step;
            if (!no_mma) {
              if constexpr (PW) {   // one tap into Z1; the first chunk overwrites (its weights are zeros when it has no block)
                if (leader && (j == 0 || (p.masks[g * kWsMaxChunks + j] & 1u)))
                  mma_i8_ss(z1, (static_cast<uint64_t>(I.a_hi) << 32) | wl, (static_cast<uint64_t>(I.b_hi) << 32) | xl, I.idesc, j == 0 ? 0u : 1u);
              } else if (j == 0) {
                ws_issue_chunk<true, alias>(I, z1, wl, xl, 0x1FFu, has_ds);
              } else {
                const uint32_t mask = p.masks[g * kWsMaxChunks + j];
                const bool ds = has_ds && (p.masks2[g * kWsMaxChunks2 + j] & 1u);
                ws_issue_chunk<false, alias>(I, z1, wl, xl, mask, ds);
              }
            }
            if (leader) mma_commit(&a_empty[as]);
            if (++as == a_slots) { as = 0; aph ^= 1u; }
            if (!w_resident) {
              if (leader) mma_commit(&w_empty[ws]);
              if (++ws == w_slots) { ws = 0; wph ^= 1u; }
            }
          }
        }
        if (leader) mma_commit(&acc_full[ab]);
        if (tl && leader && n == 0) tl[6] = clock64();                  // 6: first item issued
      }
      if (tl && leader) { tl[7] = clock64(); tl[12] = n; }              // 7: issuer done, 12: items
    }
    __syncwarp();
    tc_fence_before();
  } else if (warp >= kWsWarpLoad) {
    // =================================================================== loaders: activation tiles by LDGSTS
    // A staged tile is [R+2 input rows][32 channels][P pixels] with the 8-channel x P-byte swizzle atoms the B descriptor
    // names.  TMA tensor tiles deliver exactly this layout too, but their rows are 16-64 bytes and the TMA unit retires
    // about one row per 4 cycles (measured: 685 cycles per 8 KB stage); 16-byte LDGSTS copies issued by two warps are
    // several times faster and zero-fill the padding just the same.
    griddep_wait();          // the activations are the previous kernel's output
    const int lt = static_cast<int>(threadIdx.x) - kWsWarpLoad * 32;           // 0..63
    if (tl && lt == 0) tl[8] = clock64();                              // 8: loaders past griddepcontrol.wait
    if (!TWIN && p.use_tma) {
      // One TMA tensor tile per stage: rows / columns / channels outside the image are zero-filled by the unit (that is the
      // convolution padding).  The unit retires about one box row (P bytes) per 4 cycles, whatever the SM's warps are doing:
      // for 64-byte rows (ResNet-18 layer1: 128 rows per 8 KB stage) that is faster than the two LDGSTS warps, whose ~100
      // instructions per stage compete with the epilogue warps for issue slots; for 16 / 32-byte rows it is slower.
      if (warp == kWsWarpLoad && elect_one()) {
        uint32_t as = 0, aph = 0;
        const uint32_t a_slots_ = static_cast<uint32_t>(p.a_slots), a_stage_ = static_cast<uint32_t>(p.a_stage_bytes);
        const uint32_t box_bytes = static_cast<uint32_t>(p.a_box_bytes);
        for (uint32_t it = item0; it < n_items; it += item_step) {
          const uint32_t n_sub = (dual && 2u * it + 1u < n_tiles) ? 2u : 1u;
          for (uint32_t sub = 0; sub < n_sub; ++sub) {
            const uint32_t tt = dual ? 2u * it + sub : it;
            const uint32_t img = fdiv(tt, p.d_tpi);
            const int y0 = static_cast<int>(tt - img * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
            const int ybase = p.stride * y0 - p.ypad;
            for (uint32_t j = 0; j < n_chunks; ++j) {
              mbar_wait(&a_empty[as], aph ^ 1u);
              mbar_arrive_expect_tx(&a_full[as], box_bytes);
              tma_load_4d(a_addr + as * a_stage_, &L.tmap, 0, static_cast<int>(j * kWsCk), ybase, static_cast<int>(img), &a_full[as]);
              if (++as == a_slots_) { as = 0; aph ^= 1u; }
            }
          }
        }
      }
      __syncwarp();
    } else {
    const int x16s = p.P >> 4, rows = p.rows_in;
    const int n_ops = kWsCk * rows * (p.twin ? 2 : x16s);
    constexpr int kOps = TWIN ? kWsLoadOps + 1 : kWsLoadOps;      // twin tiles: 576 eight-byte copies per stage
    uint32_t soff[kOps];
    int32_t goff[kOps], yrow[kOps], nbytes[kOps];
#pragma unroll
    for (int k = 0; k < kOps; ++k) {
      const int o = lt + kWsLoadThreads * k;
      // twin: xq = which image of the pair (8 bytes each inside the 16-byte row), else the 16-byte column of the row
      // twin: 16 consecutive lanes take the two images of eight consecutive channels of one row = 128 contiguous bytes of the
      // stage (row-fastest lanes put every store of a warp on the same four banks: 3.5 M bank conflicts per launch under ncu)
      const int tq = o >> 1;
      const int xq = TWIN ? (o & 1) : o % x16s, y = TWIN ? (tq >> 3) % rows : (o / x16s) % rows,
                c = TWIN ? ((tq & 7) | (((tq >> 3) / rows) << 3)) : o / (x16s * rows);
      uint32_t so = static_cast<uint32_t>(y) * p.row_stride + static_cast<uint32_t>(c * p.P + xq * (TWIN ? 8 : 16));
      if (p.b_layout == 4u) so ^= ((so >> 7) & 3u) << 4;           // 64-byte swizzle
      else if (p.b_layout == 6u) so ^= ((so >> 7) & 1u) << 4;      // 32-byte swizzle
      soff[k] = so;
      goff[k] = TWIN ? (c * p.H + y) * p.in_pitch + xq * (p.C * p.H * p.in_pitch)
                       : (c * p.H + y) * p.in_pitch + xq * 16;       // + (y0 - 1) * pitch >= 0 whenever the row is inside the image
      yrow[k] = y;
      nbytes[k] = o < n_ops ? (TWIN ? min(8, p.W) : max(0, min(16, p.W - xq * 16))) : -1;
      if (TWIN) yrow[k] |= xq << 16;                                // image B of the pair may not exist (odd batch)
    }
    uint32_t as = 0, aph = 0, st_l = 0;
    // The two loader warps are latency-bound on their own instruction stream (tools/ws_timeline.py, loads only, layer1: 433 +
    // 277 cycles to issue the two stages of a tile, 712 cycles of per-tile set-up between tiles): everything a tile needs is
    // kept in registers, and a tile whose staged rows all lie inside the image (all but the first and last of an image) takes
    // the per-copy offsets and sizes as they were computed once per launch.
    // pin(): an opaque register copy.  Left alone, ptxas re-reads these kernel parameters from the constant bank inside the tile
    // loop (LDC / LDCU), and the parameter block is 4 KB of masks that the issuer walks - those loads miss the constant cache.
    const int stride_ = pin(p.stride), ypad_ = pin(p.ypad), H_ = pin(p.H), pitch_ = pin(p.in_pitch), R_ = pin(p.R);
    const uint32_t tpi_ = pin(static_cast<uint32_t>(p.tiles_per_image)), B_ = pin(static_cast<uint32_t>(p.B));
    const uint32_t a_slots_ = pin(static_cast<uint32_t>(p.a_slots)), a_stage_ = pin(static_cast<uint32_t>(p.a_stage_bytes));
    FastDiv d_tpi_ = p.d_tpi;
    d_tpi_.d = pin(d_tpi_.d); d_tpi_.mul = pin(d_tpi_.mul); d_tpi_.shr = pin(d_tpi_.shr);
    const int8_t* const x_ = reinterpret_cast<const int8_t*>(pin64(reinterpret_cast<uint64_t>(p.x)));
    const int64_t image_bytes = static_cast<int64_t>(pin64(static_cast<uint64_t>(static_cast<int64_t>(p.C) * H_ * pitch_)));
    const int64_t chunk_stride = static_cast<int64_t>(pin64(static_cast<uint64_t>(static_cast<int64_t>(kWsCk) * H_ * pitch_)));
    for (uint32_t it = item0; it < n_items; it += item_step) {
      const uint32_t n_sub = (dual && 2u * it + 1u < n_tiles) ? 2u : 1u;
      for (uint32_t sub = 0; sub < n_sub; ++sub) {
        const uint32_t tt = dual ? 2u * it + sub : it;
        const uint32_t ti = fdiv(tt, d_tpi_);                       // image (twin: image pair) of the tile
        const uint32_t img = TWIN ? 2u * ti : ti;
        const int y0 = static_cast<int>(tt - ti * tpi_) * R_;
        const bool have_b = img + 1u < B_;
        const int ybase = stride_ * y0 - ypad_;                     // first staged input row of the tile
        const int tile_off = ybase * pitch_;                        // >= 0 whenever that row is inside the image
        const bool interior = ybase >= 0 && ybase + rows <= H_ && (!TWIN || have_b);       // CTA-uniform
        // per tile: which of this thread's copies read an input row inside the image (the others zero-fill: 0 bytes
        // from offset 0), so that the per-stage loop is one LDGSTS and one address add per copy
        uint32_t go[kOps];
        int nb[kOps];
        if (interior) {
#pragma unroll
          for (int k = 0; k < kOps; ++k) { go[k] = static_cast<uint32_t>(goff[k]); nb[k] = nbytes[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < kOps; ++k) {
            const int yy = ybase + (yrow[k] & 0xffff);
            const bool ok = yy >= 0 && yy < H_ && nbytes[k] > 0 && ((yrow[k] >> 16) == 0 || have_b);
            nb[k] = ok ? nbytes[k] : min(nbytes[k], 0);
            go[k] = ok ? static_cast<uint32_t>(goff[k] + tile_off) : 0u;
          }
        }
        const int8_t* src0 = x_ + static_cast<int64_t>(img) * image_bytes + (interior ? tile_off : 0);
        for (uint32_t j = 0; j < n_chunks; ++j) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (tl && lt == 0 && st_l >= 32 && st_l < 64) tl[16 + (st_l - 32)] = clock64();     // loader got the slot of stage st_l
          const uint32_t dst0 = a_addr + as * a_stage_;
#pragma unroll
          for (int k = 0; k < kOps; ++k)
            if (nb[k] >= 0) {
              if constexpr (TWIN) cp_async8_zfill_s(dst0 + soff[k], src0 + go[k], nb[k]);
              else cp_async16_zfill_s(dst0 + soff[k], src0 + go[k], nb[k]);
            }
          src0 += chunk_stride;
          // arrives once this thread's copies have landed; the issuer orders them against tcgen05.mma with one
          // fence.proxy.async per stage after its mbarrier wait (a fence in each of the 64 loader threads measured ~1000
          // cycles per stage in round 1; the single consumer-side fence is what the PTX memory model asks for)
          cp_async_mbar_arrive(&a_full[as]);
          if (tl && lt == 0 && st_l >= 32 && st_l < 64) tl[96 + (st_l - 32)] = clock64();     // loader issued stage st_l
          ++st_l;
          if (++as == a_slots_) { as = 0; aph ^= 1u; }
        }
      }
    }
    if (tl && lt == 0) tl[9] = clock64();                              // 9: loaders done
    }
  }
  else {
    // =================================================================== weight loader (bulk copies)
    if (elect_one()) {
      uint32_t ws = 0, wph = 0;
      const uint8_t* wsrc = p.wblob + static_cast<size_t>(g) * n_chunks * static_cast<size_t>(p.w_src_stride);
      const uint32_t my_items = item0 < n_items ? (n_items - item0 + item_step - 1) / item_step : 0u;
      uint32_t passes = 0;                       // resident: one pass over the chunks; streamed: one per pixel tile
      if (p.w_resident) passes = my_items ? 1u : 0u;
      else
        for (uint32_t it = item0; it < n_items; it += item_step) passes += (dual && 2u * it + 1u < n_tiles) ? 2u : 1u;
      for (uint32_t ps = 0; ps < passes; ++ps)
        for (uint32_t j = 0; j < n_chunks; ++j) {
          uint32_t wslot;
          if (p.w_resident) {
            wslot = j;
          } else {
            wslot = ws;
            mbar_wait(&w_empty[ws], wph ^ 1u);
            if (++ws == static_cast<uint32_t>(p.w_slots)) { ws = 0; wph ^= 1u; }
          }
          uint64_t* bar = &w_full[wslot];
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(p.w_chunk_bytes));
          const uint8_t* src = wsrc + static_cast<size_t>(j) * static_cast<size_t>(p.w_src_stride);
          uint8_t* dst = smem + kWsSmemBar + wslot * static_cast<uint32_t>(p.w_chunk_bytes);
          if constexpr (PW) {
            bulk_g2s(dst, src, kWsTapBytes, bar);
          } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) bulk_g2s(dst + i * (kWsChunkBytes / 3), src + i * (kWsChunkBytes / 3), kWsChunkBytes / 3, bar);
          }
          if (p.has_ds)
            bulk_g2s(dst + kWsChunkBytes, p.wblob2 + (static_cast<size_t>(g) * n_chunks + j) * kWsTapBytes, kWsTapBytes, bar);
        }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (tl && threadIdx.x == 0) tl[10] = clock64();                      // 10: all roles done
  if (warp == kWsWarpIssue) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_base, 512);
  }
  if (tl && threadIdx.x == 0) { tl[11] = clock64(); long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); tl[13] = g; }
}

}  // namespace accel
