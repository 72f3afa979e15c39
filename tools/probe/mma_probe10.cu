// Probe 10: where does a cta_group::1 M=64 kind::i8 MMA put its 64 accumulator rows?  A = 64 rows x 32 k (K-major),
// row r = all (r+1); B = 16 x 32 of ones -> D[r][n] = 32*(r+1).  Issued twice: D lane field 0 and 16.  Reads all 128 lanes.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

__global__ void __launch_bounds__(128, 1) probe(int32_t* out, int lane_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 64 x 32 = 2 KB canonical K-major
  uint8_t* sB = smem + 4096;     // 16 x 32
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) { const int row = (i >> 8) * 8 + ((i >> 4) & 7); sA[i] = (uint8_t)(row + 1); }
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sB[i] = 1;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  for (int c = 0; c < 32; c += 4) tmem_st4(tmem_base + ((warp * 32u) << 16) + c, 0u, 0u, 0u, 0u);
  tmem_st_wait(); tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    mma_i8_ss(tmem_base + (static_cast<uint32_t>(lane_off) << 16), smem_desc_kmajor(smem_u32(sA), 128, 256),
              smem_desc_kmajor(smem_u32(sB), 128, 256), idesc_i8(64, 16), 1u);
    mma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem_base + ((warp * 32u) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x] = (int)v[0];
  out[128 + threadIdx.x] = (int)v[15];
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 32);
}

int main(int argc, char** argv) {
  const int lane_off = argc > 1 ? atoi(argv[1]) : 0;
  int32_t* d; cudaMalloc(&d, 256 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 128, 16384>>>(d, lane_off);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int32_t> h(256); cudaMemcpy(h.data(), d, 1024, cudaMemcpyDeviceToHost);
  printf("lane_off %d: accumulator row (value/32 - 1) held by each TMEM lane, -1 = untouched\n", lane_off);
  for (int l = 0; l < 128; ++l) { printf("%3d", h[l] / 32 - 1); if (l % 32 == 31) printf("\n"); }
  return 0;
}
