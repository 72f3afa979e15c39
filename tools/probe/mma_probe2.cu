// Probe 2: can several warps issue tcgen05.mma concurrently and beat the ~45-cycle single-issuer floor?
// Also: M=64 cost, and SS with a wide LBO (pairing two non-adjacent K tiles).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;
constexpr int TS = 2064, NT = 18;

__global__ void __launch_bounds__(256, 1) probe(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sx = smem;
  uint8_t* sw = smem + NT * TS + 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NT * TS + 1024 + 8192 + 1024);   // 8 barriers
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < NT * TS + 1024 + 8192; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t x_addr = smem_u32(sx), w_addr = smem_u32(sw);
  int cfg = 0;
  for (int nw : {1, 2, 4, 8}) {
    for (int variant = 0; variant < 4; ++variant) {   // 0: SS M128 N16, 1: TS M128 N16, 2: SS M64 N16, 3: SS M128 N16 wide-LBO
      __syncthreads();
      long long t0 = clock64();
      if (warp < nw && lane == 0) {
        const uint32_t M = variant == 2 ? 64 : 128;
        const uint32_t idesc = idesc_i8(M, 16);
        const uint64_t bdesc = smem_desc_kmajor(w_addr + warp * 512, 256, 128);
        const uint32_t lbo = variant == 3 ? 5 * TS : TS;
        const int per = reps / nw;
        for (int r = 0; r < per; r += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t d = tmem_base + warp * 32 + (i & 1) * 16;
            if (variant == 1) mma_i8_ts(d, tmem_base + 256 + i * 4, bdesc, idesc, 1u);
            else mma_i8_ss(d, smem_desc_kmajor(x_addr + i * TS, lbo, 128), bdesc, idesc, 1u);
          }
        }
        mma_commit(&bar[warp]);
        mbar_wait(&bar[warp], (cfg & 1));
      }
      __syncthreads();
      long long t1 = clock64();
      if (threadIdx.x == 0) out[cfg] = t1 - t0;
      ++cfg;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 8192;
  long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
  const int smem = NT * TS + 1024 + 8192 + 1024 + 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 256, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* vn[4] = {"SS M128 N16", "TS M128 N16", "SS M64 N16", "SS M128 N16 LBO=5 tiles"};
  int cfg = 0;
  for (int nw : {1, 2, 4, 8}) for (int v = 0; v < 4; ++v, ++cfg)
    printf("warps=%d %-26s %7.2f cyc/MMA (aggregate)\n", nw, vn[v], (double)h[cfg] / reps);
  return 0;
}
