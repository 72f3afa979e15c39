// Developer probe: tcgen05.ld throughput per SM.  W warps (4 or 8; warp w reads lane quadrant w % 4) read the whole 512-column
// TMEM allocation round and round with .32x32b.x16 / .x32 loads, one wait per load (D = 1) or per two loads (D = 2).
// Prints bytes per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

template <int X, int D>
__global__ void __launch_bounds__(256, 1) probe(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc_dyn(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll 1
    for (int c = 0; c < 512; c += X * D) {
      if constexpr (X == 16) {
        uint32_t a[16], b[16];
        tmem_ld16(base + c, a);
        if constexpr (D == 2) tmem_ld16(base + c + 16, b);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc ^= a[e];
        if constexpr (D == 2) {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc ^= b[e];
        }
      } else {
        uint32_t a[32], b[32];
        tmem_ld32(base + c, a);
        if constexpr (D == 2) tmem_ld32(base + c + 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) acc ^= a[e];
        if constexpr (D == 2) {
#pragma unroll
          for (int e = 0; e < 32; ++e) acc ^= b[e];
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_dyn(slot, 512); }
}

template <int X, int D>
void run(int warps, const char* name) {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 200;
  probe<X, D><<<148, warps * 32>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int i = 0; i < 148; ++i) cyc += h[i]; cyc /= 148;
  const double bytes = double(iters) * warps * 32 * 512 * 4;
  printf("%-22s warps=%d  %s  %.1f B/clk/SM  (%.0f cycles per 512-column sweep of one warp)\n", name, warps, cudaGetErrorString(e),
         bytes / cyc, cyc / iters);
  cudaFree(out); cudaFree(sink);
}

int main() {
  run<16, 1>(4, "x16, wait each");   run<16, 1>(8, "x16, wait each");
  run<16, 2>(4, "x16 x2, one wait"); run<16, 2>(8, "x16 x2, one wait");
  run<32, 1>(4, "x32, wait each");   run<32, 1>(8, "x32, wait each");
  run<32, 2>(4, "x32 x2, one wait"); run<32, 2>(8, "x32 x2, one wait");
  return 0;
}
