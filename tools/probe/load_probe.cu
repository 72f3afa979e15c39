// Developer probe: how fast can one persistent CTA per SM stage activation tiles (8 KB per stage, 16-deep ring) from an
// L2/HBM-resident int8 tensor, as a function of the SIZE OF THE CONTIGUOUS PIECES it asks for?
//   v0  LDGSTS 16 B, NCHW rows of 64 B (what conv_ws does today: 32 channels x 4 rows x 64 B per stage)
//   v1  cp.async.bulk 256 B per channel (4 contiguous rows of one channel), 32 per stage            [NCHW]
//   v2  cp.async.bulk 2 KB per image row (32 channels x 64 B contiguous), 4 per stage                [NHCW: n, y, c, x]
//   v3  LDGSTS 16 B on the NHCW layout (a warp covers 512 contiguous bytes)
//   v4  cp.async.bulk 8 KB per stage (4 rows x 32 channels contiguous)                               [N, y-tile-major]
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_probe load_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

constexpr int kSlots = 16, kStage = 8192, C = 64, H = 56, P = 64, CK = 32, R = 2, ROWS = R + 2;

template <int V>
__global__ void __launch_bounds__(192, 1) probe(const int8_t* x, int B, int n_tiles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], empty[kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(&full[s], V == 7 ? 128 : ((V == 0 || V == 3 || V == 5 || V == 6) ? 64 : 1)); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_per_image = H / R;
  if (warp == 0 && lane == 0) {                      // consumer: release every stage as soon as it is full
    uint32_t s = 0, ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int j = 0; j < C / CK; ++j) {
        mbar_wait(&full[s], ph);
        if constexpr (V == 5) { tc_fence_after(); mma_commit(&empty[s]); }       // release through tcgen05.commit, as the MMA issuer does
        else mbar_arrive(&empty[s]);
        if (++s == kSlots) { s = 0; ph ^= 1; }
      }
  } else if (warp >= 2 && (V == 7 || warp < 4)) {    // producers: warps 2, 3 (v7: 2..5)
    const int lt = threadIdx.x - 64;
    uint32_t s = 0, ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int img = t / tiles_per_image, y0 = (t % tiles_per_image) * R - 1;
      for (int j = 0; j < C / CK; ++j) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* dst = smem + s * kStage;
        if constexpr (V == 0 || V == 5 || V == 6) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int o = lt + 64 * k, xq = o & 3, y = (o >> 2) & 3, c = o >> 4;
            const int yy = y0 + y;
            const bool ok = yy >= 0 && yy < H;
            const int8_t* src = x + ((static_cast<size_t>(img) * C + j * CK + c) * H + (ok ? yy : 0)) * P + xq * 16;
            // V == 6: the last 16-byte piece of every 56-pixel row is a PARTIAL copy (8 bytes + 8 zero-filled), as conv_ws does
            const int nbytes = V == 6 ? (ok ? (xq == 3 ? 8 : 16) : 0) : (ok ? 16 : 0);
            cp_async16_zfill_s(smem_u32(dst + (y * CK + c) * P + xq * 16), src, nbytes);
          }
          cp_async_mbar_arrive(&full[s]);
        } else if constexpr (V == 7) {            // v0 with FOUR producer warps (128 threads x 4 copies): more copies in flight per SM?
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int o = lt + 128 * k, xq = o & 3, y = (o >> 2) & 3, c = o >> 4;
            const int yy = y0 + y;
            const bool ok = yy >= 0 && yy < H;
            const int8_t* src = x + ((static_cast<size_t>(img) * C + j * CK + c) * H + (ok ? yy : 0)) * P + xq * 16;
            cp_async16_zfill_s(smem_u32(dst + (y * CK + c) * P + xq * 16), src, ok ? 16 : 0);
          }
          cp_async_mbar_arrive(&full[s]);
        } else if constexpr (V == 3) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int o = lt + 64 * k, y = o >> 7, rem = o & 127;       // 128 16-byte pieces per row of 32 channels
            const int yy = y0 + y;
            const bool ok = yy >= 0 && yy < H;
            const int8_t* src = x + ((static_cast<size_t>(img) * H + (ok ? yy : 0)) * C + j * CK) * P + rem * 16;
            cp_async16_zfill_s(smem_u32(dst + y * CK * P + rem * 16), src, ok ? 16 : 0);
          }
          cp_async_mbar_arrive(&full[s]);
        } else if (lt == 0) {
          if constexpr (V == 1) {
            const int ya = max(y0, 0), yb = min(y0 + ROWS, H);
            mbar_arrive_expect_tx(&full[s], CK * (yb - ya) * P);
            for (int c = 0; c < CK; ++c)
              bulk_g2s(dst + c * ROWS * P, x + ((static_cast<size_t>(img) * C + j * CK + c) * H + ya) * P, (yb - ya) * P, &full[s]);
          } else if constexpr (V == 2) {
            const int ya = max(y0, 0), yb = min(y0 + ROWS, H);
            mbar_arrive_expect_tx(&full[s], CK * (yb - ya) * P);
            for (int yy = ya; yy < yb; ++yy)
              bulk_g2s(dst + (yy - y0) * CK * P, x + ((static_cast<size_t>(img) * H + yy) * C + j * CK) * P, CK * P, &full[s]);
          } else {
            mbar_arrive_expect_tx(&full[s], kStage);
            const size_t off = (static_cast<size_t>(t) * (C / CK) + j) * kStage % (static_cast<size_t>(B) * C * H * P - kStage);
            bulk_g2s(dst, x + (off & ~size_t(15)), kStage, &full[s]);
          }
        }
        if (++s == kSlots) { s = 0; ph ^= 1; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && sink) atomicAdd(sink, smem[0]);
}

// Handshake probe: the loader / consumer pipeline of v5 plus the accumulator hand-over of conv_ws: every ITEM (4 stages) the
// consumer commits to acc_full[set], 8 "epilogue" warps wait for it and arrive on acc_empty[set], and the consumer must see
// acc_empty before it starts the item after next (two sets).  MODE 0: suspended try_wait everywhere (mbar_wait), 1: the epilogue
// warps spin, 2: epilogue warps and consumer spin.
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe_hs(const int8_t* x, int B, int n_tiles, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kSlots], empty[kSlots], acc_full[2], acc_empty[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(&full[s], 64); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_per_image = H / R;
  constexpr int kStagesPerItem = 4;                  // two tiles x two chunks
  const int n_items = n_tiles / 2;
  if (warp < 8) {                                    // "epilogue"
    uint32_t n = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
      const uint32_t ab = n & 1;
      if (MODE >= 1) mbar_wait_spin(&acc_full[ab], (n >> 1) & 1); else mbar_wait(&acc_full[ab], (n >> 1) & 1);
      tc_fence_after();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
    }
  } else if (warp == 8) {
    if (lane == 0) {
      uint32_t s = 0, ph = 0, n = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
        const uint32_t ab = n & 1;
        if (MODE >= 2) mbar_wait_spin(&acc_empty[ab], ((n >> 1) & 1) ^ 1); else mbar_wait(&acc_empty[ab], ((n >> 1) & 1) ^ 1);
        for (int j = 0; j < kStagesPerItem; ++j) {
          if (MODE >= 2) mbar_wait_spin(&full[s], ph); else mbar_wait(&full[s], ph);
          tc_fence_after();
          mma_commit(&empty[s]);
          if (++s == kSlots) { s = 0; ph ^= 1; }
        }
        mma_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= 10) {
    const int lt = threadIdx.x - 320;
    uint32_t s = 0, ph = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x)
      for (int sub = 0; sub < 2; ++sub) {
        const int t = 2 * it + sub;
        const int img = t / tiles_per_image, y0 = (t % tiles_per_image) * R - 1;
        for (int j = 0; j < C / CK; ++j) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* dst = smem + s * kStage;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int o = lt + 64 * k, xq = o & 3, y = (o >> 2) & 3, c = o >> 4;
            const int yy = y0 + y;
            const bool ok = yy >= 0 && yy < H;
            const int8_t* src = x + ((static_cast<size_t>(img) * C + j * CK + c) * H + (ok ? yy : 0)) * P + xq * 16;
            cp_async16_zfill_s(smem_u32(dst + (y * CK + c) * P + xq * 16), src, ok ? 16 : 0);
          }
          cp_async_mbar_arrive(&full[s]);
          if (++s == kSlots) { s = 0; ph ^= 1; }
        }
      }
  }
  __syncthreads();
  if (threadIdx.x == 0 && sink) atomicAdd(sink, smem[0]);
}

template <int MODE>
void run_hs(const int8_t* x, int B, unsigned long long* sink, const char* name) {
  const int n_tiles = B * (H / R);
  cudaFuncSetAttribute(probe_hs<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kStage);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe_hs<MODE><<<148, 384, kSlots * kStage>>>(x, B, n_tiles, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) probe_hs<MODE><<<148, 384, kSlots * kStage>>>(x, B, n_tiles, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = static_cast<double>(n_tiles) * (C / CK) * kStage;
  printf("%-44s %8.1f us per pass  %7.2f TB/s staged  (%s)\n", name, ms * 100.0, bytes / (ms / 10 * 1e-3) / 1e12, cudaGetErrorString(err));
}


// v8: a transcription of the conv_ws_kernel loader (per-launch tables soff / goff / yrow / nbytes, per-tile go / nb, 384 threads,
// loader = warps 10 and 11, consumer = one thread of warp 8 that releases through tcgen05.commit, 8 idle "epilogue" warps
// that wait on a barrier nobody completes until the end) with the geometry of ResNet-18 layer1 in a parameter struct.
struct V8Params { int C, H, W, B, P, R, rows_in, stride, ypad, in_pitch, a_slots, a_stage_bytes, n_chunks, dual, tiles_per_image, n_tiles; unsigned row_stride, b_layout; const int8_t* x; };
template <int VAR>
__global__ void __launch_bounds__(384, 1) probe_v8(const __grid_constant__ V8Params p, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t a_full[kSlots], a_empty[kSlots], never;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(&a_full[s], 64); mbar_init(&a_empty[s], 1); }
    mbar_init(&never, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t a_addr = smem_u32(smem);
  const uint32_t n_tiles = p.n_tiles, dual = p.dual, n_chunks = p.n_chunks;
  const uint32_t n_items = dual ? (n_tiles + 1u) >> 1 : n_tiles;
  const uint32_t item0 = blockIdx.x, item_step = gridDim.x;
  if (warp == 8) {
    if (lane == 0) {
      uint32_t as = 0, aph = 0;
      for (uint32_t it = item0; it < n_items; it += item_step) {
        const uint32_t n_sub = (dual && 2u * it + 1u < n_tiles) ? 2u : 1u;
        for (uint32_t sub = 0; sub < n_sub; ++sub)
          for (uint32_t j = 0; j < n_chunks; ++j) {
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            mma_commit(&a_empty[as]);
            if (++as == static_cast<uint32_t>(p.a_slots)) { as = 0; aph ^= 1u; }
          }
      }
      mbar_arrive(&never);
    }
  } else if (warp < 8) {
    if (VAR & 1) mbar_wait(&never, 0);            // idle epilogue warps asleep on a barrier
  } else if (warp >= 10) {
    const int lt = static_cast<int>(threadIdx.x) - 320;
    const int x16s = p.P >> 4, rows = p.rows_in;
    const int n_ops = 32 * rows * x16s;
    constexpr int kOps = 8;
    uint32_t soff[kOps];
    int32_t goff[kOps], yrow[kOps], nbytes[kOps];
#pragma unroll
    for (int k = 0; k < kOps; ++k) {
      const int o = lt + 64 * k;
      const int xq = o % x16s, y = (o / x16s) % rows, c = o / (x16s * rows);
      uint32_t so = static_cast<uint32_t>(y) * p.row_stride + static_cast<uint32_t>(c * p.P + xq * 16);
      if (!(VAR & 4)) { if (p.b_layout == 4u) so ^= ((so >> 7) & 3u) << 4;
      else if (p.b_layout == 6u) so ^= ((so >> 7) & 1u) << 4; }
      soff[k] = so;
      goff[k] = (c * p.H + y) * p.in_pitch + xq * 16;
      yrow[k] = y;
      nbytes[k] = o < n_ops ? max(0, min(16, p.W - xq * 16)) : -1;
    }
    uint32_t as = 0, aph = 0;
    const int64_t chunk_stride = static_cast<int64_t>(32) * p.H * p.in_pitch;
    for (uint32_t it = item0; it < n_items; it += item_step) {
      const uint32_t n_sub = (VAR & 16) ? 2u : ((dual && 2u * it + 1u < n_tiles) ? 2u : 1u);
#pragma unroll
      for (uint32_t sub = 0; sub < n_sub; ++sub) {
        const uint32_t tt = dual ? 2u * it + sub : it;
        const uint32_t ti = (VAR & 32) ? tt / 28u : tt / static_cast<uint32_t>(p.tiles_per_image);
        const uint32_t img = ti;
        const int y0 = static_cast<int>(tt - ti * static_cast<uint32_t>(p.tiles_per_image)) * p.R;
        uint32_t go[kOps];
        int nb[kOps];
#pragma unroll
        for (int k = 0; k < kOps; ++k) {
          const int yy = p.stride * y0 - p.ypad + (yrow[k] & 0xffff);
          const bool ok = yy >= 0 && yy < p.H && nbytes[k] > 0;
          nb[k] = ok ? nbytes[k] : min(nbytes[k], 0);
          go[k] = ok ? static_cast<uint32_t>(goff[k] + (p.stride * y0 - p.ypad) * p.in_pitch) : 0u;
          if (VAR & 2) { nb[k] = ok ? 16 : 0; }                    // VAR bit 1: sizes 0 / 16 only (no partial copies)
        }
        const int8_t* src0 = p.x + static_cast<int64_t>(img) * p.C * p.H * p.in_pitch;
        const uint32_t n_ch = (VAR & 16) ? 2u : n_chunks;
#pragma unroll
        for (uint32_t j = 0; j < n_ch; ++j) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          const uint32_t dst0 = a_addr + as * static_cast<uint32_t>(p.a_stage_bytes);
          if (VAR & 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int o = lt + 64 * k, xq = o & 3, y = (o >> 2) & 3, c = o >> 4;
              const int yy = y0 - 1 + y;
              const bool ok = yy >= 0 && yy < H;
              const int8_t* src = p.x + ((static_cast<size_t>(img) * C + j * CK + c) * H + (ok ? yy : 0)) * P + xq * 16;
              cp_async16_zfill_s(dst0 + (y * CK + c) * P + xq * 16, src, ok ? 16 : 0);
            }
          } else {
#pragma unroll
          for (int k = 0; k < kOps; ++k)
            if (nb[k] >= 0) cp_async16_zfill_s(dst0 + soff[k], src0 + go[k], nb[k]);
          }
          src0 += chunk_stride;
          cp_async_mbar_arrive(&a_full[as]);
          if (++as == static_cast<uint32_t>(p.a_slots)) { as = 0; aph ^= 1u; }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && sink) atomicAdd(sink, smem[0]);
}
template <int VAR>
void run_v8(const int8_t* x, int B, unsigned long long* sink, const char* name) {
  V8Params p{};
  p.C = C; p.H = H; p.W = 56; p.B = B; p.P = P; p.R = R; p.rows_in = ROWS; p.stride = 1; p.ypad = 1; p.in_pitch = P;
  p.a_slots = kSlots; p.a_stage_bytes = kStage; p.n_chunks = C / CK; p.dual = 1; p.tiles_per_image = H / R; p.n_tiles = B * (H / R);
  p.row_stride = CK * P; p.b_layout = 4u; p.x = x;
  cudaFuncSetAttribute(probe_v8<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kStage);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe_v8<VAR><<<148, 384, kSlots * kStage>>>(p, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) probe_v8<VAR><<<148, 384, kSlots * kStage>>>(p, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = static_cast<double>(p.n_tiles) * (C / CK) * kStage;
  printf("%-44s %8.1f us per pass  %7.2f TB/s staged  (%s)\n", name, ms * 100.0, bytes / (ms / 10 * 1e-3) / 1e12, cudaGetErrorString(err));
}

int g_extra_smem = 0;      // extra dynamic shared memory (carve-out experiment: less L1 left for the copies in flight)
template <int V>
void run(const int8_t* x, int B, unsigned long long* sink, const char* name) {
  const int n_tiles = B * (H / R);
  cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kStage + g_extra_smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) probe<V><<<148, V == 7 ? 192 : 128, kSlots * kStage + g_extra_smem>>>(x, B, n_tiles, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) probe<V><<<148, V == 7 ? 192 : 128, kSlots * kStage + g_extra_smem>>>(x, B, n_tiles, sink);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = static_cast<double>(n_tiles) * (C / CK) * kStage;
  printf("%-44s %8.1f us per pass  %7.2f TB/s staged  (%s)\n", name, ms * 100.0, bytes / (ms / 10 * 1e-3) / 1e12, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 256;
  g_extra_smem = argc > 2 ? atoi(argv[2]) : 0;
  printf("extra smem %d\n", g_extra_smem);
  const size_t n = static_cast<size_t>(B) * C * H * P;
  int8_t* x; unsigned long long* sink;
  cudaMalloc(&x, n); cudaMalloc(&sink, 8);
  cudaMemset(x, 1, n); cudaMemset(sink, 0, 8);
  printf("tensor %.1f MB, stages of %d B, ring %d\n", n / 1e6, kStage, kSlots);
  run<0>(x, B, sink, "v0 LDGSTS16 NCHW (64-B rows)");
  run<1>(x, B, sink, "v1 bulk 256 B per channel (NCHW)");
  run<2>(x, B, sink, "v2 bulk 2 KB per image row (NHCW)");
  run<3>(x, B, sink, "v3 LDGSTS16 NHCW (2-KB rows)");
  run<4>(x, B, sink, "v4 bulk 8 KB per stage");
  run<5>(x, B, sink, "v5 = v0, slots released by tcgen05.commit");
  run<6>(x, B, sink, "v6 = v0 with partial (8-byte) zero-fill copies");
  run<7>(x, B, sink, "v7 = v0 with four producer warps");
  run_v8<0>(x, B, sink, "v8 conv_ws loader transcription");
  run_v8<1>(x, B, sink, "v8 + idle warps asleep on a barrier");
  run_v8<2>(x, B, sink, "v8 without partial copies");
  run_v8<4>(x, B, sink, "v8 without the swizzle");
  run_v8<8>(x, B, sink, "v8 with inline addresses (as v0)");
  run_v8<12>(x, B, sink, "v8 inline + no swizzle");
  run_v8<16>(x, B, sink, "v8 tables, loops unrolled (2 x 2 stages)");
  run_v8<32>(x, B, sink, "v8 tables, constant division");
  run_v8<48>(x, B, sink, "v8 tables, unrolled + constant division");
  run_v8<56>(x, B, sink, "v8 inline, unrolled + constant division");
  run_hs<0>(x, B, sink, "hs0 accumulator hand-over, suspended waits");
  run_hs<1>(x, B, sink, "hs1 epilogue warps spin");
  run_hs<2>(x, B, sink, "hs2 epilogue warps + consumer spin");
  return 0;
}
