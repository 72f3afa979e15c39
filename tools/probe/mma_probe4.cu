// Probe 4: lean, fully warp-uniform issue path (operands live in uniform registers): what is the real
// cadence of small-N TS-mode tcgen05.mma kind::i8, per issuing warp and in aggregate?
#include <cstdio>
#include <cstdlib>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

template <int N>
__device__ __forceinline__ void run(int nw, int warp, uint32_t tmem_base, uint32_t w_addr, uint64_t* bar, int reps,
                                    int& phase, const uint16_t* meta, int use_meta) {
  if (warp < nw) {
    const uint32_t idesc = idesc_i8(128, N);
    const uint64_t bdesc0 = smem_desc_kmajor(w_addr + warp * 8192, N * 16, 128);
    const int per = reps / nw;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < per; r += 32) {
      uint32_t pk = 0;
      if (use_meta) pk = meta[(r + lane) & 1023];     // lane l holds the packed (D col | A col << 9) word of op l
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint32_t d, a;
        if (use_meta) {
          const uint32_t v = __shfl_sync(0xffffffffu, pk, i);
          d = tmem_base + (v & 0x1ffu);
          a = tmem_base + (v >> 9);
        } else {
          d = tmem_base + ((warp * 2 + (i & 1)) * N) % 384;
          a = tmem_base + 448 + (i & 7) * 4;
        }
        const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(((i & 15) * 512) >> 4);
        if (elect_one()) mma_i8_ts(d, a, bdesc, idesc, 1u);
      }
    }
    if (elect_one()) mma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], phase & 1);
    ++phase;
  }
}

__global__ void __launch_bounds__(256, 1) probe(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sw = smem;                       // 64 KB of B tiles
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  uint16_t* meta = reinterpret_cast<uint16_t*>(smem + 65536 + 128);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) meta[i] = (uint16_t)(((i * 5) % 11) * 16 | ((448 + (i % 8) * 4) << 9));
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t w_addr = smem_u32(sw);
  int cfg = 0;
  int phase = 0;
  for (int um = 0; um < 2; ++um)
  for (int N : {16, 32}) {
    for (int nw : {1, 2, 4, 8}) {
      __syncthreads();
      long long t0 = clock64();
      if (N == 16) run<16>(nw, warp, tmem_base, w_addr, bar, reps, phase, meta, um);
      else run<32>(nw, warp, tmem_base, w_addr, bar, reps, phase, meta, um);
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
  const int smem = 65536 + 128 + 2048 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 256, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int cfg = 0;
  for (int um = 0; um < 2; ++um) for (int N : {16, 32}) for (int nw : {1, 2, 4, 8}) { 
    printf("%s TS M128 N=%3d warps=%d  %7.2f cyc/MMA  (floor %d)\n", um ? "meta/shfl" : "computed ", N, nw, (double)h[cfg] / reps, N / 2); ++cfg; }
  return 0;
}
