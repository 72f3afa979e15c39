// Micro-benchmark + semantics probe for tcgen05.mma kind::i8 on sm_100a (developer tool, not product code).
//   * issue/execute rate of small-N MMAs, SS (A from smem) vs TS (A from TMEM), same vs rotating accumulators
//   * TMEM A-operand layout check: A written with tcgen05.st 32x32b must reproduce the SS result
//   * TS A-operand column alignment (start at +1 column = +4 bytes of K)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

constexpr int TS = 2064;  // K-tile stride in smem (as in the product kernel)
constexpr int NT = 18;

struct Res { long long cyc[16]; int diff[8]; };

template <int N, bool kTS, bool kRotate>
__device__ long long run(uint32_t tmem_base, uint32_t a_col, uint32_t x_addr, uint32_t w_addr, uint64_t* bar, uint32_t& phase, int reps) {
  const uint32_t idesc = idesc_i8(128, N);
  const uint64_t adesc0 = smem_desc_kmajor(x_addr, TS, 128);
  const uint64_t bdesc0 = smem_desc_kmajor(w_addr, N * 16, 128);
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; r += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t d = tmem_base + (kRotate ? (i * N) % 256 : 0);
      const uint32_t win = i;                      // window start tile (K slot pair i, i+1)
      if (kTS) mma_i8_ts(d, tmem_base + a_col + win * 4, bdesc0, idesc, 1u);
      else mma_i8_ss(d, adesc0 + ((win * TS) >> 4), bdesc0, idesc, 1u);
    }
  }
  long long t1 = clock64();
  mma_commit(bar);
  mbar_wait(bar, phase); phase ^= 1;
  long long t2 = clock64();
  return ((t1 - t0) << 32) | (t2 - t0);
}

__global__ void __launch_bounds__(160, 1) probe(const int8_t* X /*[128][NT*16]*/, const int8_t* Wt /*[256][32]*/, Res* out, int reps, int max_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sx = smem;                       // NT * TS
  uint8_t* sw = smem + NT * TS + 1024;      // up to 256 x 32 B tile (canonical K-major: [2][N][16])
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NT * TS + 1024 + 8192 + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 4) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  // X -> canonical K-major smem tiles; also into TMEM columns [256 .. 256 + NT*4) via tcgen05.st (lane = row)
  if (warp < 4) {
    const int r = threadIdx.x;
    for (int t = 0; t < NT; ++t) {
      const uint4 v = *reinterpret_cast<const uint4*>(X + (r * NT + t) * 16);
      *reinterpret_cast<uint4*>(sx + t * TS + r * 16) = v;
    }
  }
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) {
    // W tile rows n (0..255), 32 K bytes; canonical [k_half][n][16] with LBO = N*16 chosen at run time -> store for N=256
    sw[i] = 0;
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (warp < 4) {
    const int r = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int t = 0; t < NT; ++t) {
      const uint4 v = *reinterpret_cast<const uint4*>(X + (r * NT + t) * 16);
      tmem_st4(tmem_base + lane_base + 256 + t * 4, v.x, v.y, v.z, v.w);
    }
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();

  uint32_t phase = 0;
  auto fill_w = [&](int N) {   // all threads: W tile for this N in canonical layout, LBO = N*16
    __syncthreads();
    for (int i = threadIdx.x; i < N * 32; i += blockDim.x) {
      const int n = i / 32, k = i % 32;
      sw[(k / 16) * (N * 16) + n * 16 + (k % 16)] = (uint8_t)Wt[n * 32 + k];
    }
    fence_proxy_async_smem();
    __syncthreads();
  };
  auto zero_acc = [&]() {      // clear accumulator columns 0..255 through an MMA-free path: tcgen05.st zeros
    if (warp < 4) {
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
      for (int c = 0; c < 256; c += 4) tmem_st4(tmem_base + lane_base + c, 0, 0, 0, 0);
      tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
  };
  const uint32_t x_addr = smem_u32(sx), w_addr = smem_u32(sw);
  long long r_;
#define RUN(idx, N, TSM, ROT)                                                         \
  fill_w(N); zero_acc();                                                              \
  if (warp == 4 && lane == 0) { r_ = run<N, TSM, ROT>(tmem_base, 256, x_addr, w_addr, bar, phase, reps); out->cyc[idx] = r_; } \
  __syncthreads();
  RUN(0, 16, false, true)
  RUN(1, 16, false, false)
  RUN(2, 16, true, true)
  RUN(3, 16, true, false)
  RUN(4, 64, false, true)
  RUN(5, 64, true, true)
  RUN(6, 128, false, true)
  RUN(7, 128, true, true)
  RUN(8, 256, false, false)
  RUN(9, 256, true, false)
  RUN(10, 32, false, true)
  RUN(11, 32, true, true)

  // ---- semantics: one MMA, N=16, window 3: SS vs TS vs TS with A shifted by one column (+4 K bytes)
  uint32_t v[16];
  for (int mode = 0; mode < max_mode; ++mode) {
    fill_w(16); zero_acc();
    if (warp == 4 && lane == 0) {
      const uint32_t idesc = idesc_i8(128, 16);
      const uint64_t bdesc = smem_desc_kmajor(w_addr, 256, 128);
      if (mode == 0) mma_i8_ss(tmem_base, smem_desc_kmajor(x_addr + 3 * TS, TS, 128), bdesc, idesc, 0u);
      if (mode == 1) mma_i8_ts(tmem_base, tmem_base + 256 + 3 * 4, bdesc, idesc, 0u);
      if (mode == 2) mma_i8_ts(tmem_base, tmem_base + 256 + 3 * 4 + 1, bdesc, idesc, 0u);
      mma_commit(bar); mbar_wait(bar, phase); phase ^= 1;
    }
    __syncthreads(); tc_fence_after();
    if (warp < 4) {
      tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16), v);
      tmem_ld_wait();
      // expected on the fly
      const int r = threadIdx.x;
      int bad = 0;
      for (int n = 0; n < 16; ++n) {
        int acc = 0;
        for (int k = 0; k < 32; ++k) {
          const int kk = 3 * 16 + k + (mode == 2 ? 4 : 0);
          acc += (int)X[r * NT * 16 + kk] * (int)Wt[n * 32 + k];
        }
        if (acc != (int)v[n]) ++bad;
      }
      if (bad) atomicAdd(&out->diff[mode], bad);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
  }
  if (warp == 4) tmem_dealloc_dyn(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 4096;
  std::vector<int8_t> X(128 * NT * 16), W(256 * 32);
  srand(1);
  for (auto& v : X) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : W) v = (int8_t)(rand() % 256 - 128);
  int8_t *dX, *dW; Res* dR;
  cudaMalloc(&dX, X.size()); cudaMalloc(&dW, W.size()); cudaMalloc(&dR, sizeof(Res));
  cudaMemcpy(dX, X.data(), X.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size(), cudaMemcpyHostToDevice);
  cudaMemset(dR, 0, sizeof(Res));
  const int smem = NT * TS + 1024 + 8192 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int pass = 0; pass < 3; ++pass) {
    const int grid = pass == 1 ? 148 : 1;
    const int max_mode = pass == 2 ? 3 : 2;
    probe<<<grid, 160, smem>>>(dX, dW, dR, reps, max_mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    Res r; cudaMemcpy(&r, dR, sizeof(Res), cudaMemcpyDeviceToHost);
    const char* names[12] = {"SS N=16 rot", "SS N=16 same", "TS N=16 rot", "TS N=16 same", "SS N=64 rot", "TS N=64 rot",
                             "SS N=128 rot", "TS N=128 rot", "SS N=256", "TS N=256", "SS N=32 rot", "TS N=32 rot"};
    printf("grid=%d reps=%d max_mode=%d\n", grid, reps, max_mode);
    for (int i = 0; i < 12; ++i)
      printf("  %-14s issue %7.1f cyc/MMA   total %7.1f cyc/MMA\n", names[i], (double)(r.cyc[i] >> 32) / reps,
             (double)(r.cyc[i] & 0xffffffffLL) / reps);
    printf("  semantics: SS mismatches=%d  TS mismatches=%d  TS(+1 col) mismatches=%d (of 2048)\n", r.diff[0], r.diff[1], r.diff[2]);
    cudaMemset(dR, 0, sizeof(Res));
  }
  return 0;
}
