// Probe 8: which shared-memory byte does tcgen05.mma kind::i8 read for A[m][k] when A is MN-major with 128-byte
// swizzle?  A is filled with its own byte offset (two passes: low 7 bits, then bits 7..11), B is the identity,
// so D[m][k] = A[m][k] = offset.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

__host__ __device__ constexpr uint32_t idesc_i8_amn(uint32_t M, uint32_t Nn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((Nn >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint64_t smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(int32_t* out, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 8 KB
  uint8_t* sB = smem + 8192;                // 32 x 32 B identity, canonical K-major no swizzle
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192 + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int rg = i >> 8, ks = (i >> 7) & 1, r8 = (i >> 4) & 7, col = i & 15;
    sB[i] = (rg * 8 + r8 == ks * 16 + col) ? 1 : 0;
  }
  if (warp == 0) { tmem_alloc_dyn(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  for (int pass = 0; pass < 2; ++pass) {
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sA[i] = pass == 0 ? (i & 0x7f) : ((i >> 7) & 0x3f);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      mma_i8_ss(tmem_base, smem_desc_sw(smem_u32(sA), lbo, sbo, layout), smem_desc_kmajor(smem_u32(sB), 128, 256),
                idesc_i8_amn(128, 32), 0u);
      mma_commit(&bar[0]);
      mbar_wait(&bar[0], pass);
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < 32; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((warp * 32u) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) out[(pass * 128 + threadIdx.x) * 32 + c0 + i] = static_cast<int32_t>(v[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 64);
}

int main(int argc, char** argv) {
  const uint32_t lbo = argc > 1 ? atoi(argv[1]) : 1024, sbo = argc > 2 ? atoi(argv[2]) : 1024, layout = argc > 3 ? atoi(argv[3]) : 2;
  int32_t* d; cudaMalloc(&d, 2 * 128 * 32 * 4);
  const int smem = 8192 + 1024 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(d, lbo, sbo, layout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int32_t> h(2 * 128 * 32);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  printf("lbo %u sbo %u layout %u: smem byte offset read for A[m][k]\n", lbo, sbo, layout);
  for (int m : {0, 1, 2, 15, 16, 17, 31, 32, 63, 64, 100, 127}) {
    printf("m=%3d:", m);
    for (int k = 0; k < 32; ++k) printf(" %4d", h[m * 32 + k] + 128 * h[(128 + m) * 32 + k]);
    printf("\n");
  }
  // compare against the TMA SWIZZLE_128B image of rows k (128 B each): off = k*128 + ((m/16) ^ (k%8))*16 + m%16
  long bad = 0;
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) {
    const int want = (k / 8) * (int)sbo + (k % 8) * 128 + (((m / 16) ^ (k % 8)) * 16) + m % 16;
    if (want != h[m * 32 + k] + 128 * h[(128 + m) * 32 + k]) ++bad;
  }
  printf("mismatches vs 'row k = 128 contiguous pixels, TMA 128B swizzle': %ld of 4096\n", bad);
  return 0;
}
