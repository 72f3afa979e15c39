// Probe 5: bitmask-driven issue loop (operands derived by uniform arithmetic from a per-window row mask).
#include <cstdio>
#include <cstdlib>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

// One chunk: masks[w] (w = 0..7) has bit g set when block-row g has an op in window w.  Ops are stored
// window-major / row-minor.  Issuer `me` of `nIss` takes the windows w with w % nIss == me; pre[w] is the
// op ordinal of the first op of window w (prefix popcount, precomputed by the host).
template <int kIss>
__device__ __forceinline__ void issue_chunk(const uint4 mw, const uint2 pre, uint32_t acc0, uint32_t xa0, uint64_t bdesc0,
                                            uint32_t idesc, int me, uint32_t leader) {
  const uint32_t mws[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if ((w % kIss) != me) continue;
    uint32_t m = (mws[w >> 1] >> ((w & 1) * 16)) & 0xffffu;
    const uint32_t p = ((w < 4 ? pre.x : pre.y) >> ((w & 3) * 8)) & 0xffu;
    uint64_t bdesc = bdesc0 + static_cast<uint64_t>(p * 32u);
    while (m) {
      const uint32_t g = __ffs(m) - 1;
      m &= m - 1;
      if (leader) mma_i8_ts(acc0 + g * 16, xa0 + w * 4, bdesc, idesc, 1u);
      bdesc += 32;
    }
  }
}

__global__ void __launch_bounds__(256, 1) probe(long long* out, int reps, uint32_t mask16) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sw = smem;                       // 64 KB of B tiles
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  uint32_t* meta = reinterpret_cast<uint32_t*>(smem + 65536 + 128);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) smem[i] = (uint8_t)(i * 7 + 3);
  const int per_w = __popc(mask16);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) meta[i] = mask16 | (mask16 << 16);
    uint32_t a = 0, b = 0;
    for (int w = 0; w < 4; ++w) { a |= (uint32_t)(w * per_w) << (8 * w); b |= (uint32_t)((w + 4) * per_w) << (8 * w); }
    meta[4] = a; meta[5] = b;
  }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  const uint32_t w_addr = smem_u32(sw);
  const uint32_t idesc = idesc_i8(128, 16);
  int cfg = 0, phase = 0;
  for (int nw : {1, 2, 4}) {
    __syncthreads();
    long long t0 = clock64();
    if (warp < nw) {
      const uint32_t leader = elect_one();
      const int chunks = reps / (8 * per_w);
      for (int c = 0; c < chunks; ++c) {
        uint4 mw = *reinterpret_cast<const uint4*>(meta);
        uint2 pre = *reinterpret_cast<const uint2*>(meta + 4);
        mw.x = __shfl_sync(0xffffffffu, mw.x, 0); mw.y = __shfl_sync(0xffffffffu, mw.y, 0);
        mw.z = __shfl_sync(0xffffffffu, mw.z, 0); mw.w = __shfl_sync(0xffffffffu, mw.w, 0);
        pre.x = __shfl_sync(0xffffffffu, pre.x, 0); pre.y = __shfl_sync(0xffffffffu, pre.y, 0);
        const uint64_t bdesc0 = smem_desc_kmajor(w_addr, 256, 128);
        const uint32_t xa0 = tmem_base + 176 + (c & 1) * 36;
        if (nw == 1) issue_chunk<1>(mw, pre, tmem_base, xa0, bdesc0, idesc, warp, leader);
        else if (nw == 2) issue_chunk<2>(mw, pre, tmem_base, xa0, bdesc0, idesc, warp, leader);
        else issue_chunk<4>(mw, pre, tmem_base, xa0, bdesc0, idesc, warp, leader);
      }
      if (leader) mma_commit(&bar[warp]);
      __syncwarp();
      mbar_wait(&bar[warp], phase & 1);
      ++phase;
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[cfg] = t1 - t0;
    ++cfg;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 8192;
  long long* d; cudaMalloc(&d, 64 * 8);
  const int smem = 65536 + 128 + 2048 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (uint32_t mask : {0x7ffu, 0x155u, 0x021u}) {
    cudaMemset(d, 0, 64 * 8);
    probe<<<1, 256, smem>>>(d, reps, mask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const int per_w = __builtin_popcount(mask);
    const int issued = reps / (8 * per_w) * 8 * per_w;
    int cfg = 0;
    for (int nw : {1, 2, 4}) { printf("mask=%03x (%2d rows/window) issuers=%d  %7.2f cyc/MMA\n", mask, per_w, nw, (double)h[cfg] / issued); ++cfg; }
  }
  return 0;
}
