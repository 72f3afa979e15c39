// Probe 3: TS-mode (A in TMEM) tcgen05.mma kind::i8 cadence vs N and number of issuing warps.
#include <cstdio>
#include <cstdlib>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

__global__ void __launch_bounds__(256, 1) probe(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sw = smem;                       // 64 KB of B tiles
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t w_addr = smem_u32(sw);
  int cfg = 0;
  int phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int N : {16, 32, 64, 128, 256}) {
    for (int nw : {1, 2, 4}) {
      for (int dist = 0; dist < 2; ++dist) {   // 0: alternate 2 accumulators per warp, 1: rotate over 8 accumulators (N<=32)
        __syncthreads();
        long long t0 = clock64();
        if (warp < nw && lane == 0) {
          const uint32_t idesc = idesc_i8(128, N);
          // B tile: N rows x 32 bytes, K-major core matrices: LBO = N*16 (between K halves), SBO = 128
          const uint64_t bdesc = smem_desc_kmajor(w_addr + warp * 8192, N * 16, 128);
          const int per = reps / nw;
          for (int r = 0; r < per; r += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t d;
              if (N <= 32) d = tmem_base + (dist ? ((warp * 8 + i) * N) % 384 : (warp * 2 + (i & 1)) * N);
              else if (N == 64) d = tmem_base + ((warp * 2 + (i & 1)) * 64) % 384;
              else if (N == 128) d = tmem_base + ((warp + i) & 1) * 128;
              else d = tmem_base;
              mma_i8_ts(d, tmem_base + 448 + (i & 7) * 4, bdesc, idesc, 1u);
            }
          }
          mma_commit(&bar[warp]);
          mbar_wait(&bar[warp], phase[warp] & 1);
        }
        if (warp < nw) ++phase[warp];
        __syncthreads();
        long long t1 = clock64();
        if (threadIdx.x == 0) out[cfg] = t1 - t0;
        ++cfg;
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 8192;
  long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
  const int smem = 65536 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 256, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int cfg = 0;
  for (int N : {16, 32, 64, 128, 256}) for (int nw : {1, 2, 4}) for (int dist = 0; dist < 2; ++dist, ++cfg)
    printf("TS M128 N=%3d warps=%d dist=%d  %7.2f cyc/MMA  (floor %d)\n", N, nw, dist, (double)h[cfg] / reps, N / 2);
  return 0;
}
