// Probe 7: A operand straight from TMA.  A [32 channels][2 rows x 64 pixels] box of an NCHW int8 tensor, written by
// TMA with SWIZZLE_128B, is an MN-major (pixel-contiguous) K=32 x M=128 tile: can tcgen05.mma kind::i8 consume it
// as-is (idesc a_major = MN), is the result right for shifted / out-of-bounds boxes, and what does one MMA cost?
// Also: does a traversal stride of 2 (elementStrides) give the stride-2 convolution gather?
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include "../../resnet_accel_b200/csrc/ptx.cuh"
using namespace accel;

constexpr int C = 32, H = 8, W = 56, PITCH = 64, N = 64;

__host__ __device__ constexpr uint32_t idesc_i8_amn(uint32_t M, uint32_t Nn) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((Nn >> 3) << 17) | ((M >> 4) << 24);
}
// MN-major, 128-byte swizzle: 8 k-rows x 128 B atoms; sbo = byte stride between 8-row k groups
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}

struct Args {
  CUtensorMap tm1, tm2;
  const uint8_t* wts;   // canonical K-major no-swizzle B tile, N x 32
  int32_t* out;         // [3][128][N]
  long long* cyc;
  uint8_t* dump;
  int x0, y0, reps, mask, amajor;
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 8 chunks x 4 KB
  uint8_t* sB = smem + 8 * 4096;            // 256 x 32 B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * 4096 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sB[i] = a.wts[i % (N * 32)];
  for (int i = threadIdx.x; i < 8 * 4096; i += blockDim.x) sA[i] = 0x55;
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc_dyn(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = *slot;
  uint32_t par[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int test = 0; test < ((a.mask & 1) ? ((a.mask & 8) ? 1 : (a.mask & 16) ? 2 : 3) : 0); ++test) {
    // test 0: box at (x0, y0), stride-1 map.  test 1: box at (-1, -1): left / top padding.  test 2: stride-2 map at (-1, y0)
    if (threadIdx.x == 0) {
      if (test < 2) {
        mbar_arrive_expect_tx(&bar[0], 4096);
        tma_load_4d(smem_u32(sA), &a.tm1, test == 0 ? a.x0 : -1, test == 0 ? a.y0 : -1, 0, 0, &bar[0]);
      } else {
        mbar_arrive_expect_tx(&bar[0], 4096);
        tma_load_4d(smem_u32(sA), &a.tm2, -1, a.y0, 0, 0, &bar[0]);
      }
      mbar_wait(&bar[0], par[0]); par[0] ^= 1;
      if (test == 0) for (int i = 0; i < 4096; ++i) a.dump[i] = sA[i];
      tc_fence_after();
      mma_i8_ss(tmem_base, smem_desc_mn_sw128(smem_u32(sA), 1024, 1024), smem_desc_kmajor(smem_u32(sB), 128, 256),
                (a.amajor ? idesc_i8_amn(128, N) : idesc_i8(128, N)), 0u);
      mma_commit(&bar[1]);
      mbar_wait(&bar[1], par[1]); par[1] ^= 1;
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((warp * 32u) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) a.out[(test * 128 + threadIdx.x) * N + c0 + i] = static_cast<int32_t>(v[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  // ---- cadence: reps MMAs cycling through the 8 A chunks, for several N
  int cfg = 0;
  for (int nn : {16, 32, 64, 96, 128, 192, 256}) {
    for (int variant = 0; variant < 2; ++variant) {   // 0: A MN-major SW128 from smem, 1: A K-major no-swizzle from smem
      __syncthreads();
      const long long t0 = clock64();
      if (threadIdx.x == 0 && (a.mask & (2 << variant))) {
        const uint32_t idesc = variant == 0 ? idesc_i8_amn(128, nn) : idesc_i8(128, nn);
        const uint64_t bdesc = smem_desc_kmajor(smem_u32(sB), 128, 256);
        for (int r = 0; r < a.reps; r += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint64_t ad = variant == 0 ? smem_desc_mn_sw128(smem_u32(sA) + i * 4096, 1024, 1024)
                                             : smem_desc_kmajor(smem_u32(sA) + i * 4096, 128, 256);
            mma_i8_ss(tmem_base + (i & 1) * 256, ad, bdesc, idesc, 1u);
          }
        }
        mma_commit(&bar[2]);
        mbar_wait(&bar[2], par[2]); par[2] ^= 1;
      }
      __syncthreads();
      const long long t1 = clock64();
      if (threadIdx.x == 0) a.cyc[cfg] = t1 - t0;
      ++cfg;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc_dyn(tmem_base, 512);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 4096;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  std::vector<int8_t> x(C * H * PITCH), wt(N * 32);
  srand(1);
  for (size_t i = 0; i < x.size(); ++i) { int xx = i % PITCH, y = (i / PITCH) % H, c = i / (PITCH * H); x[i] = static_cast<int8_t>((xx & 63) | ((y & 1) << 6) | ((c & 1) << 7)); }
  for (auto& v : wt) v = static_cast<int8_t>(rand() % 256 - 128);
  std::vector<uint8_t> wcan(N * 32);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 32; ++k) wcan[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + k % 16] = static_cast<uint8_t>(wt[n * 32 + k]);
  uint8_t *dx, *dw; int32_t* dout; long long* dc;
  cudaMalloc(&dx, x.size()); cudaMalloc(&dw, wcan.size()); cudaMalloc(&dout, 3 * 128 * N * 4); cudaMalloc(&dc, 64 * 8);
  cudaMemcpy(dx, x.data(), x.size(), cudaMemcpyHostToDevice); cudaMemcpy(dw, wcan.data(), wcan.size(), cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, 3 * 128 * N * 4); cudaMemset(dc, 0, 64 * 8);
  Args a;
  memset(&a, 0, sizeof(a));
  cuuint64_t gd[4] = {W, H, C, 1}, gs[3] = {PITCH, PITCH * H, (cuuint64_t)PITCH * H * C};
  cuuint32_t bx[4] = {64, 2, 32, 1}, es[4] = {1, 1, 1, 1};
  CUresult r1 = enc(&a.tm1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, dx, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  // stride 2 in x and y: 32 output pixels x 4 output rows per 128-lane tile
  cuuint32_t bx2[4] = {64, 8, 32, 1}, es2[4] = {2, 2, 1, 1};
  CUresult r2 = enc(&a.tm2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, dx, gd, gs, bx2, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: stride1 %d stride2 %d\n", (int)r1, (int)r2);
  if (r1 != CUDA_SUCCESS) return 1;
  if (r2 != CUDA_SUCCESS) a.tm2 = a.tm1;
  uint8_t* ddump; cudaMalloc(&ddump, 4096); cudaMemset(ddump, 0xEE, 4096); a.dump = ddump; a.wts = dw; a.out = dout; a.cyc = dc; a.x0 = argc > 4 ? atoi(argv[4]) : 1; a.y0 = 3; a.reps = reps; a.mask = argc > 2 ? atoi(argv[2]) : 7; a.amajor = argc > 3 ? atoi(argv[3]) : 1;
  const int smem = 8 * 4096 + 8192 + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  { std::vector<uint8_t> dmp(4096); cudaMemcpy(dmp.data(), ddump, 4096, cudaMemcpyDeviceToHost);
    for (int r = 0; r < 10; ++r) { printf("row %2d:", r); for (int ch = 0; ch < 8; ++ch) { int b = dmp[r * 128 + ch * 16]; printf("  c%d y%d x%2d", b >> 7, (b >> 6) & 1, b & 63); } printf("\n"); } }
  std::vector<int32_t> out(3 * 128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  auto X = [&](int c, int y, int xx) -> int { return (y < 0 || y >= H || xx < 0 || xx >= W) ? 0 : x[(c * H + y) * PITCH + xx]; };
  for (int test = 0; test < 3; ++test) {
    long bad = 0, first = -1;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        long ref = 0;
        for (int c = 0; c < 32; ++c) {
          int v;
          if (test == 0) v = X(c, a.y0 + m / 64, a.x0 + m % 64);
          else if (test == 1) v = X(c, -1 + m / 64, -1 + m % 64);
          else v = X(c, a.y0 + 2 * (m / 32), -1 + 2 * (m % 32));
          ref += v * wt[n * 32 + c];
        }
        if (ref != out[(test * 128 + m) * N + n]) { if (first < 0) first = m * N + n; ++bad; }
      }
    printf("test %d: %ld mismatches of %d (first at m=%ld n=%ld)\n", test, bad, 128 * N, first / N, first % N);
  }
  long long h[64]; cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
  int cfg = 0;
  for (int nn : {16, 32, 64, 96, 128, 192, 256})
    for (int v = 0; v < 2; ++v, ++cfg)
      printf("N=%3d %-22s %7.2f cyc/MMA\n", nn, v == 0 ? "A MN-major SW128 (SS)" : "A K-major no-swz (SS)", (double)h[cfg] / reps);
  return 0;
}
