// Probe 9: weight-stationary orientation for the 3x3 stride-1 convolution.
//   D[co][px] (TMEM lane = output channel, column = pixel)  +=  W_tap[co][32 ch] (A, K-major, smem)  x  X[32 ch][px] (B, MN-major)
// The activation tile [32 ch][R+2 rows][pitch] arrives by ONE TMA box (dims ordered x, c, y so that an image row of a
// channel group is one swizzle atom).  Tap (kh, kw): B window = rows kh .. kh+R-1 (descriptor start + kh * LBO),
// D column base shifted by (1 - kw): all 9 taps reuse the same staged tile.  Checks the full tile against a CPU conv
// and reports cycles per MMA.  argv: pitch (64 / 32 / 16), R, y0
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

constexpr int C = 64, CO = 128;

__host__ __device__ constexpr uint32_t idesc_i8_bmn(uint32_t M, uint32_t Nn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((Nn >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

struct Args {
  CUtensorMap tm;
  const uint8_t* wts;   // [2 chunks][9 taps][4 KB]
  int32_t* out;         // [128 co][N]
  long long* cyc;
  int pitch, R, y0, H, reps, mode;
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                       // 2 x 9 x 4 KB = 72 KB
  uint8_t* sX = smem + 73728;               // 2 chunks x up to 32 x 10 rows x 64 B = 2 x 20 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 73728 + 2 * 20480);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5;
  const int N = a.R * a.pitch;
  const uint32_t act_bytes = 32u * (a.R + 2) * a.pitch;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 73728 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sW)[i] = reinterpret_cast<const uint4*>(a.wts)[i];
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t lane_base = (warp * 32u) << 16;
  for (int c = 0; c < 256; c += 4) tmem_st4(tmem_base + lane_base + c, 0u, 0u, 0u, 0u);
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t layout = a.pitch == 64 ? 4u : a.pitch == 32 ? 6u : 0u;
  // swizzled: LBO = stride between MN atoms (image rows), SBO = stride between 8-channel groups; no swizzle: the names swap
  const uint32_t row_stride = 32u * a.pitch, grp_stride = 8u * a.pitch;
  const uint32_t lbo = layout ? row_stride : grp_stride, sbo = layout ? grp_stride : row_stride;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], 2 * act_bytes);
    for (int ch = 0; ch < 2; ++ch) tma_load_4d(smem_u32(sX) + ch * 20480, &a.tm, 0, ch * 32, a.y0 - 1, 0, &bar[0]);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = idesc_i8_bmn(128, N);
    for (int ch = 0; ch < 2; ++ch)
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap % 3;
        const uint64_t ad = smem_desc_kmajor(smem_u32(sW) + (ch * 9 + tap) * 4096, 128, 256);
        const uint64_t bd = smem_desc(smem_u32(sX) + ch * 20480 + kh * row_stride, lbo, sbo, layout);
        mma_i8_ss(tmem_base + (kw == 1 ? 0 : (N - 2) + (kw == 0 ? 2 : 0)), ad, bd, idesc, 1u);
      }
    mma_commit(&bar[1]);
    mbar_wait(&bar[1], 0);
  }
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    uint32_t u[16];
    tmem_ld16(tmem_base + lane_base + c0, v);
    tmem_ld16(tmem_base + lane_base + (N - 2) + 1 + c0, u);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) a.out[threadIdx.x * N + c0 + i] = static_cast<int32_t>(v[i] + u[i]);
  }
  tc_fence_before();
  __syncthreads();
  // ---- cadence of the 18-MMA tile loop
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_i8_bmn(128, N);
    for (int r = 0; r < a.reps; ++r)
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap % 3;
          const uint64_t ad = smem_desc_kmajor(smem_u32(sW) + (ch * 9 + tap) * 4096, 128, 256);
          const uint64_t bd = smem_desc(smem_u32(sX) + ch * 20480 + kh * row_stride, lbo, sbo, layout);
          mma_i8_ss(tmem_base + (kw == 1 ? 0 : (N - 2) + (kw == 0 ? 2 : 0)), ad, bd, idesc, 1u);
        }
    mma_commit(&bar[2]);
    mbar_wait(&bar[2], 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) a.cyc[0] = clock64() - t0;
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 256);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int pitch = argc > 1 ? atoi(argv[1]) : 64, R = argc > 2 ? atoi(argv[2]) : 2, y0 = argc > 3 ? atoi(argv[3]) : 0;
  const int W = pitch - 8 + (pitch == 16 ? 6 : pitch == 32 ? 4 : 0), H = 9, N = R * pitch, reps = 256;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  std::vector<int8_t> x(C * H * pitch, 0), wt(CO * C * 9);
  srand(1);
  for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int xx = 0; xx < W; ++xx) x[(c * H + y) * pitch + xx] = static_cast<int8_t>(rand() % 256 - 128);
  for (auto& v : wt) v = static_cast<int8_t>(rand() % 256 - 128);    // wt[co][c][tap]
  std::vector<uint8_t> wcan(73728);
  for (int ch = 0; ch < 2; ++ch) for (int tap = 0; tap < 9; ++tap) for (int co = 0; co < CO; ++co) for (int k = 0; k < 32; ++k)
    wcan[(ch * 9 + tap) * 4096 + (co / 8) * 256 + (k / 16) * 128 + (co % 8) * 16 + k % 16] = static_cast<uint8_t>(wt[(co * C + ch * 32 + k) * 9 + tap]);
  uint8_t *dx, *dw; int32_t* dout; long long* dc;
  cudaMalloc(&dx, x.size()); cudaMalloc(&dw, wcan.size()); cudaMalloc(&dout, 128 * N * 4); cudaMalloc(&dc, 64);
  cudaMemcpy(dx, x.data(), x.size(), cudaMemcpyHostToDevice); cudaMemcpy(dw, wcan.data(), wcan.size(), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 128 * N * 4);
  Args a;
  memset(&a, 0, sizeof(a));
  // dims ordered (x, c, y, n): an image row of 8 channels is one swizzle atom, rows of the tile are MN atoms
  cuuint64_t gd[4] = {(cuuint64_t)pitch, C, H, 1}, gs[3] = {(cuuint64_t)pitch * H, (cuuint64_t)pitch, (cuuint64_t)pitch * H * C};
  cuuint32_t bx[4] = {(cuuint32_t)pitch, 32, (cuuint32_t)(R + 2), 1}, es[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = pitch == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : pitch == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r1 = enc(&a.tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, dx, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("pitch %d W %d R %d N %d y0 %d encode %d\n", pitch, W, R, N, y0, (int)r1);
  if (r1 != CUDA_SUCCESS) return 1;
  a.wts = dw; a.out = dout; a.cyc = dc; a.pitch = pitch; a.R = R; a.y0 = y0; a.H = H; a.reps = reps; a.mode = argc > 4 ? atoi(argv[4]) : 0;
  const int smem = 73728 + 2 * 20480 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<int32_t> out(128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  auto X = [&](int c, int y, int xx) -> int { return (y < 0 || y >= H || xx < 0 || xx >= W) ? 0 : x[(c * H + y) * pitch + xx]; };
  long bad = 0, first = -1, checked = 0;
  for (int co = 0; co < CO; ++co)
    for (int r = 0; r < R; ++r)
      for (int xx = 0; xx < W; ++xx) {
        long ref = 0;
        for (int c = 0; c < C; ++c) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
          ref += X(c, y0 + r + kh - 1, xx + kw - 1) * wt[(co * C + c) * 9 + kh * 3 + kw];
        ++checked;
        if (ref != out[co * N + r * pitch + xx]) { if (first < 0) first = co * N + r * pitch + xx; ++bad; }
      }
  printf("conv tile: %ld mismatches of %ld (first at co=%ld px=%ld)\n", bad, checked, first / N, first % N);
  long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  printf("%.1f cycles per MMA (M=128, N=%d, SS: A K-major 4 KB + B MN-major %d B)\n", (double)h / (reps * 18), N, N * 32);
  return 0;
}
