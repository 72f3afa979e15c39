// C ABI of libaccel_b200.so (declared in include/accel_b200.h).  Thin: validate, fill the kernel
// parameter block, launch on the caller's stream.  No device allocation, no CPU arithmetic.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/accel_b200.h"
#include "bsr_tc.cuh"
#include "bsr_tcp.cuh"
#include "conv_ws.cuh"
#include "gemm_ws.cuh"
#include "stem_ws.cuh"
#include "plan.h"
#include "simple_kernels.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(ACCEL_DMA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                            \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);   \
  } while (0)

constexpr int kMaxDevices = 64;
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
// SM count of the CURRENT device (a process may drive several GPUs)
int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}
// grid for grid-stride kernels: a multiple of the SM count, capped by the work
int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  int64_t g = need < cap ? need : cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

long long* g_timeline = nullptr;   // accel_debug_set_timeline
int g_dbg_flags = std::getenv("ACCEL_DBG_FLAGS") ? std::atoi(std::getenv("ACCEL_DBG_FLAGS")) : 0;   // developer aid (stage isolation)
bool g_no_persist = std::getenv("ACCEL_NO_PERSIST") != nullptr;   // developer switch: one-shot kernel everywhere
long long g_ws_launches = 0;         // accel_debug_counter(0)
bool g_ws_s2_narrow = std::getenv("ACCEL_WS_S2_NARROW") != nullptr;   // developer switch: 64-pixel stride-2 tiles even with streamed weights
bool g_no_pdl = std::getenv("ACCEL_NO_PDL") != nullptr;            // developer switch: plain stream-ordered launches
int g_ws_tma = std::getenv("ACCEL_WS_TMA") ? std::atoi(std::getenv("ACCEL_WS_TMA")) : 3;      // developer switch, conv_ws activation stages by TMA: 0 never (LDGSTS), 1 for 64-byte rows, 2 also 32-byte, 3 (default) also 16-byte rows
bool g_no_twin = std::getenv("ACCEL_NO_TWIN") != nullptr;          // developer switch: one image per tile even for 7-pixel rows
bool g_no_ws = std::getenv("ACCEL_NO_WS") != nullptr;              // developer switch: never take the weight-stationary conv path
int g_stem_rows = std::getenv("ACCEL_STEM_ROWS") ? std::atoi(std::getenv("ACCEL_STEM_ROWS")) : 0;      // developer switch: pooled rows per stem item
bool g_no_fast_epi = std::getenv("ACCEL_NO_FAST_EPI") != nullptr;  // developer switch: conversion instructions in every epilogue
bool g_no_gemm_ws = std::getenv("ACCEL_NO_GEMM_WS") != nullptr;    // developer switch: GEMMs stay on the gather kernels
int g_gemm_ws_cg = std::getenv("ACCEL_GEMM_WS_CG") ? std::atoi(std::getenv("ACCEL_GEMM_WS_CG")) : 2;   // 1: no CTA pairs
int g_gemm_ws_stages = std::getenv("ACCEL_GEMM_WS_STAGES") ? std::atoi(std::getenv("ACCEL_GEMM_WS_STAGES")) : 0;
long long g_ws_fast_launches = 0;    // accel_debug_counter(2): conv_ws launches that took the conversion-free epilogue
long long g_gw_launches = 0;         // accel_debug_counter(1)
// cudaFuncSetAttribute is per device: one flag and one status per device ordinal
std::once_flag g_attr_once_dev[kMaxDevices];
cudaError_t g_attr_err_dev[kMaxDevices] = {};
#define g_attr_once g_attr_once_dev[current_device()]
#define g_attr_err g_attr_err_dev[current_device()]
constexpr int kSmemTwoCtas = 113 * 1024;    // per CTA when two CTAs share an SM (227 KB usable, 1 KB reserved each)
constexpr int kSmemOneCta = 200 * 1024;
constexpr int kSmemPersist = 220 * 1024;    // the persistent kernel owns its SM
constexpr int kSmemWs = 227 * 1024;         // conv_ws_kernel: everything an SM has
// conv_ws_kernel<mode, residual mode, saturation counting>: one instantiation per epilogue variant (stride 2 has no residual)
// Launch with programmatic stream serialization: the grid may begin (prologue, TMEM allocation, resident weight load)
// while the previous kernel on the stream drains; the kernels order their dependent reads and all their writes behind
// griddepcontrol.wait themselves.  Captured into CUDA graphs as a programmatic dependency edge.
template <typename Params>
cudaError_t launch_overlapped(void (*kfn)(Params), unsigned grid, unsigned block, size_t smem, cudaStream_t st, const Params& arg) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kfn, arg);
}

using WsKernelFn = void (*)(accel::WsLaunch);
// conv_ws_kernel<mode, residual mode, saturation counting, conversion-free epilogue>.  The conversion-free epilogue is only
// instantiated for residual mode 4 (integer add): A/B on one box, batch 256 - residual layers 81 -> 78 us (layer1), but the
// layers without a residual lose 2-3 us to its extra FP instructions (their XU pipe was not the limiter)
template <int MODE, int RES>
WsKernelFn ws_kernel_pick(bool sat, bool fast) {
  if constexpr (RES == 4) {
    if (fast)
      return sat ? static_cast<WsKernelFn>(accel::conv_ws_kernel<MODE, RES, true, true>)
                 : static_cast<WsKernelFn>(accel::conv_ws_kernel<MODE, RES, false, true>);
  }
  return sat ? static_cast<WsKernelFn>(accel::conv_ws_kernel<MODE, RES, true, false>)
             : static_cast<WsKernelFn>(accel::conv_ws_kernel<MODE, RES, false, false>);
}
WsKernelFn ws_kernel_ptr(int mode, int resmode, bool sat, bool fast) {
  if (mode == accel::kWsModeS2) return resmode == 0 ? ws_kernel_pick<accel::kWsModeS2, 0>(sat, fast) : nullptr;
  if (mode == accel::kWsModePw || mode == accel::kWsModeTwinPw) {      // pointwise: no residual / the integer residual add / the general divide
    const bool tw = mode == accel::kWsModeTwinPw;
    switch (resmode) {
      case 0: return tw ? ws_kernel_pick<accel::kWsModeTwinPw, 0>(sat, fast) : ws_kernel_pick<accel::kWsModePw, 0>(sat, fast);
      case 1: return tw ? ws_kernel_pick<accel::kWsModeTwinPw, 1>(sat, fast) : ws_kernel_pick<accel::kWsModePw, 1>(sat, fast);
      case 2: return tw ? ws_kernel_pick<accel::kWsModeTwinPw, 2>(sat, fast) : ws_kernel_pick<accel::kWsModePw, 2>(sat, fast);
      case 3: return tw ? ws_kernel_pick<accel::kWsModeTwinPw, 3>(sat, fast) : ws_kernel_pick<accel::kWsModePw, 3>(sat, fast);
      default: return tw ? ws_kernel_pick<accel::kWsModeTwinPw, 4>(sat, fast) : ws_kernel_pick<accel::kWsModePw, 4>(sat, fast);
    }
  }
  if (mode == accel::kWsModeTwin) {
    switch (resmode) {
      case 0: return ws_kernel_pick<accel::kWsModeTwin, 0>(sat, fast);
      case 1: return ws_kernel_pick<accel::kWsModeTwin, 1>(sat, fast);
      case 2: return ws_kernel_pick<accel::kWsModeTwin, 2>(sat, fast);
      case 3: return ws_kernel_pick<accel::kWsModeTwin, 3>(sat, fast);
      default: return ws_kernel_pick<accel::kWsModeTwin, 4>(sat, fast);
    }
  }
  switch (resmode) {
    case 0: return ws_kernel_pick<accel::kWsModeS1, 0>(sat, fast);
    case 1: return ws_kernel_pick<accel::kWsModeS1, 1>(sat, fast);
    case 2: return ws_kernel_pick<accel::kWsModeS1, 2>(sat, fast);
    case 3: return ws_kernel_pick<accel::kWsModeS1, 3>(sat, fast);
    default: return ws_kernel_pick<accel::kWsModeS1, 4>(sat, fast);
  }
}
const void* ws_kernel_fn(int mode, int resmode, bool sat, bool fast) {
  return reinterpret_cast<const void*>(ws_kernel_ptr(mode, resmode, sat, fast));
}

using GwKernelFn = void (*)(accel::GwLaunch);
GwKernelFn gw_kernel_ptr(int cg, int outk) {
  if (cg == 2) return outk == 0 ? accel::gemm_ws_kernel<2, 0> : (outk == 1 ? accel::gemm_ws_kernel<2, 1> : accel::gemm_ws_kernel<2, 2>);
  return outk == 0 ? accel::gemm_ws_kernel<1, 0> : (outk == 1 ? accel::gemm_ws_kernel<1, 1> : accel::gemm_ws_kernel<1, 2>);
}
const void* gw_kernel_fn(int cg, int outk) { return reinterpret_cast<const void*>(gw_kernel_ptr(cg, outk)); }

void set_kernel_attrs() {
  const void* fns[] = {reinterpret_cast<const void*>(accel::bsr_tc_kernel<accel::kModeGemm>),
                       reinterpret_cast<const void*>(accel::bsr_tc_kernel<accel::kModeConv3>),
                       reinterpret_cast<const void*>(accel::bsr_tc_kernel<accel::kModeConv7>),
                       reinterpret_cast<const void*>(accel::bsr_tc_kernel<accel::kModeDirect>)};
  for (const void* f : fns) {
    if (g_attr_err == cudaSuccess)
      g_attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOneCta);
  }
  const void* pfns[] = {reinterpret_cast<const void*>(accel::bsr_tcp_kernel<accel::kModeGemm>),
                        reinterpret_cast<const void*>(accel::bsr_tcp_kernel<accel::kModeConv3>),
                        reinterpret_cast<const void*>(accel::bsr_tcp_kernel<accel::kModeConv7>)};
  for (const void* f : pfns) {
    if (g_attr_err == cudaSuccess)
      g_attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPersist);
  }
  for (int m = 0; m < 5; ++m)
    for (int r = 0; r < 5; ++r)
      for (int t = 0; t < 4; ++t)
        if (const void* f = ws_kernel_fn(m, r, (t & 1) != 0, (t & 2) != 0))
          if (g_attr_err == cudaSuccess) g_attr_err = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWs);
  if (g_attr_err == cudaSuccess)
    g_attr_err = cudaFuncSetAttribute(reinterpret_cast<const void*>(accel::stem_ws_kernel),
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWs);
  for (int cg = 1; cg <= 2; ++cg)
    for (int k = 0; k < 3; ++k)
      if (g_attr_err == cudaSuccess)
        g_attr_err = cudaFuncSetAttribute(gw_kernel_fn(cg, k), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWs);
}

int check_epilogue(const accel_epilogue* epi, const accel_out_layout* lay, const void* out, int32_t max_channels) {
  if (!epi || !lay || !out) return fail(ACCEL_INVALID_CONFIG, "null epilogue / layout / output");
  const int kinds = (epi->flags & ACCEL_OUT_I8 ? 1 : 0) + (epi->flags & ACCEL_OUT_I32 ? 1 : 0) +
                    (epi->flags & ACCEL_OUT_F32 ? 1 : 0);
  if (kinds != 1) return fail(ACCEL_INVALID_CONFIG, "exactly one of ACCEL_OUT_I8 / I32 / F32 must be set");
  if ((epi->flags & (ACCEL_OUT_I8 | ACCEL_OUT_F32)) && !epi->chan_scale)
    return fail(ACCEL_INVALID_CONFIG, "chan_scale required for int8 / float32 output");
  if (epi->residual && !(epi->flags & ACCEL_OUT_I8))
    return fail(ACCEL_INVALID_CONFIG, "residual add is defined on the int8 output only");
  if (epi->n_channels < 0 || epi->n_channels > max_channels)
    return fail(ACCEL_INVALID_CONFIG, "n_channels exceeds n_block_rows*14");
  if (lay->rows_per_image <= 0) return fail(ACCEL_INVALID_CONFIG, "rows_per_image must be positive");
  return ACCEL_OK;
}

// add_residual_int8 divides by s_out (golden_models.cpp:486).  The kernel may replace the IEEE divide by
//   q0 = s * rcp;  e = fma(-q0, s_out, s);  q = fma(e, rcp, q0)      with rcp = RN(1 / s_out)
// when that reproduces the reference's int8 result for EVERY (main, residual) int8 pair - decided here,
// once per scale triple, by exhaustive comparison with the true quotient.
// Returns 0: IEEE divide needed, 1: the 3-instruction sequence is exact, 2: even the single multiply  q = s * rcp  gives the
// reference's int8 for every pair (the weight-stationary kernel then drops the two FMAs as well), 3: the whole float
// sequence equals the saturating INTEGER sum main + residual for every pair (matched scales, the usual identity add).
int residual_divide_mode(float s_main, float s_res, float s_out) {
  static std::mutex mu;
  static std::map<std::array<uint32_t, 3>, int> cache;
  std::array<uint32_t, 3> key;
  std::memcpy(&key[0], &s_main, 4); std::memcpy(&key[1], &s_res, 4); std::memcpy(&key[2], &s_out, 4);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  bool ok = std::isfinite(s_out) && s_out != 0.f, ok_mul = true, ok_add = true;
  const float rcp = 1.0f / s_out;
  ok = ok && std::isfinite(rcp);
  auto sat8 = [](float f) {
    const float r = std::nearbyintf(f);
    return r > 127.f ? 127 : (r < -128.f ? -128 : static_cast<int>(r));
  };
  for (int a = -128; ok && a < 128; ++a) {
    volatile float am = static_cast<float>(a) * s_main;    // volatile: no contraction into an FMA
    for (int r = -128; r < 128; ++r) {
      volatile float rr = static_cast<float>(r) * s_res;
      volatile float sum = am + rr;
      const float s = sum;
      const float truth = s / s_out;
      volatile float q0v = s * rcp;
      const float q0 = q0v;
      const float e = std::fmaf(-q0, s_out, s);
      const float q = std::fmaf(e, rcp, q0);
      if (!std::isfinite(truth) || !std::isfinite(q) || sat8(q) != sat8(truth)) { ok = false; break; }
      if (sat8(q0) != sat8(truth)) ok_mul = false;
      const int isum = a + r;
      if (sat8(truth) != (isum > 127 ? 127 : (isum < -128 ? -128 : isum))) ok_add = false;
    }
  }
  const int mode = !ok ? 0 : (ok_add ? 3 : (ok_mul ? 2 : 1));
  cache[key] = mode;
  return mode;
}
bool residual_fast_divide_ok(float s_main, float s_res, float s_out) { return residual_divide_mode(s_main, s_res, s_out) >= 1; }

// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda needed)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// int8 tensor, dims / box innermost first, strides in bytes for dims 1..rank-1 (multiples of 16)
bool encode_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                 const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_NONE,
                 CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE>
cudaError_t launch_persistent(const accel::TcLaunch& L, unsigned items, int smem, cudaStream_t st) {
  const unsigned ctas = items < static_cast<unsigned>(sm_count()) ? items : static_cast<unsigned>(sm_count());
  accel::bsr_tcp_kernel<MODE><<<ctas, accel::kPThreads, smem, st>>>(L, items);
  return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_mode(const accel::TcLaunch& L, unsigned ctas, int smem, cudaStream_t st) {
  accel::bsr_tc_kernel<MODE><<<ctas, accel::kThreads, smem, st>>>(L);
  return cudaGetLastError();
}

// One or more launches: each carries the schedule tables of a range of block-row groups in its parameter block.
int launch_tc(const accel::Plan* P, accel::TcParams& prm, int mode, int smem, cudaStream_t st,
              const CUtensorMap* tmap = nullptr) {
  std::call_once(g_attr_once, set_kernel_attrs);
  if (g_attr_err != cudaSuccess) return cuda_fail(g_attr_err, "cudaFuncSetAttribute(smem)");
  const int n_groups = static_cast<int>(P->groups.size());
  const int64_t m_tiles = (prm.M + accel::kTileM - 1) / accel::kTileM;
  if (n_groups == 0 || m_tiles == 0) return ACCEL_OK;
  prm.blob = P->ws_dev + P->off_blob;
  {  // largest element offset the epilogue can form: (images-1)*image_stride + channels*chan_stride + one plane
    const accel_out_layout& l = prm.lay;
    const int64_t images = (prm.M + l.rows_per_image - 1) / l.rows_per_image;
    const int64_t plane = l.row_len > 0 ? (l.rows_per_image / l.row_len + 1) * l.row_pitch + l.row_len * l.row_stride
                                        : l.rows_per_image * l.row_stride;
    const int64_t span = images * l.image_stride + static_cast<int64_t>(prm.epi.n_channels + 14) * l.chan_stride + plane;
    prm.out_small = span >= 0 && span < (1ll << 31) && l.chan_stride < (1ll << 27);
  }
  prm.timeline = g_timeline;
  prm.dbg_flags = g_dbg_flags;
  if (prm.epi.residual) {
    prm.res_fast = residual_fast_divide_ok(prm.epi.res_scale_main, prm.epi.res_scale_res, prm.epi.res_scale_out) ? 1 : 0;
    prm.res_rcp = 1.0f / prm.epi.res_scale_out;
  }
  // TMA-fed launches run the persistent kernel (one CTA per SM, every role keeps going across tiles)
  bool persistent = prm.use_tma && tmap && mode != accel::kModeDirect && !g_no_persist;
  int psmem = 0;
  if (persistent) {
    const int avail = kSmemPersist - accel::kPSmemRing - accel::kRingSlack;
    int slots = avail / prm.slot_bytes;
    if (slots > accel::kPMaxRingSlots) slots = accel::kPMaxRingSlots;
    if (slots < 2) persistent = false;
    else { prm.ring_slots = slots; psmem = accel::kPSmemRing + slots * prm.slot_bytes + accel::kRingSlack; }
  }
  static thread_local accel::TcLaunch L;    // 28 KB: keep it off the stack
  int g0 = 0;
  while (g0 < n_groups) {
    int g1 = g0;
    uint32_t nb = 0, no = 0;
    while (g1 < n_groups && g1 - g0 < accel::kMaxGroupsL) {
      const accel::GroupRec& G = P->groups[g1];
      const uint32_t gb = G.batch_end - G.batch_begin;
      const uint32_t go = (g1 + 1 < n_groups ? P->groups[g1 + 1].op_begin : static_cast<uint32_t>(P->ops.size())) - G.op_begin;
      if (nb + gb > accel::kMaxBatchesL || no + go > accel::kMaxOpsL) break;
      nb += gb; no += go; ++g1;
    }
    if (g1 == g0) return fail(ACCEL_INVALID_CONFIG, "block-row group exceeds the launch tables");
    const uint32_t b0 = P->groups[g0].batch_begin, o0 = P->groups[g0].op_begin;
    L.p = prm;
    L.p.d_wo = accel::make_fastdiv(static_cast<uint32_t>(prm.Wo));
    L.p.d_ho = accel::make_fastdiv(static_cast<uint32_t>(prm.Ho));
    L.p.d_groups = accel::make_fastdiv(static_cast<uint32_t>(g1 - g0));
    L.p.d_rpi = accel::make_fastdiv(static_cast<uint32_t>(prm.lay.rows_per_image < (1ll << 31) ? prm.lay.rows_per_image : 0));
    L.p.d_rowlen = accel::make_fastdiv(static_cast<uint32_t>(prm.lay.row_len));
    if (tmap) L.tmap = *tmap;
    L.n_groups = static_cast<uint32_t>(g1 - g0);
    for (int g = g0; g < g1; ++g) {
      accel::GroupRec G = P->groups[g];
      G.batch_begin -= b0; G.batch_end -= b0; G.op_begin -= o0;
      L.groups[g - g0] = G;
    }
    if (nb) std::memcpy(L.batches, P->batches.data() + b0, nb * sizeof(uint32_t));
    if (no) std::memcpy(L.ops, P->ops.data() + o0, no * sizeof(accel::OpRec));
    const int64_t ctas = m_tiles * (g1 - g0);
    if (ctas > INT_MAX) return fail(ACCEL_INVALID_CONFIG, "grid too large");
    cudaError_t e;
    if (persistent) {
      switch (mode) {
        case accel::kModeGemm: e = launch_persistent<accel::kModeGemm>(L, static_cast<unsigned>(ctas), psmem, st); break;
        case accel::kModeConv3: e = launch_persistent<accel::kModeConv3>(L, static_cast<unsigned>(ctas), psmem, st); break;
        default: e = launch_persistent<accel::kModeConv7>(L, static_cast<unsigned>(ctas), psmem, st); break;
      }
    } else
    switch (mode) {
      case accel::kModeGemm: e = launch_mode<accel::kModeGemm>(L, static_cast<unsigned>(ctas), smem, st); break;
      case accel::kModeConv3: e = launch_mode<accel::kModeConv3>(L, static_cast<unsigned>(ctas), smem, st); break;
      case accel::kModeConv7: e = launch_mode<accel::kModeConv7>(L, static_cast<unsigned>(ctas), smem, st); break;
      default: e = launch_mode<accel::kModeDirect>(L, static_cast<unsigned>(ctas), smem, st); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "bsr_tc_kernel launch");
    g0 = g1;
  }
  return ACCEL_OK;
}

}  // namespace

// weight-stationary convolution state of a plan (conv_ws.cuh): one prepared geometry
struct WsState {
  bool ready = false;
  int32_t c_in = 0, c_out = 0, taps = 0, n_chunks = 0, n_groups = 0;
  const uint8_t* blob = nullptr;
  uint16_t masks[accel::kWsMaxGroups * accel::kWsMaxChunks] = {};
};

// dense-equivalent GEMM state of a plan (gemm_ws.cuh)
struct GwState {
  bool ready = false;
  int64_t n_pad = 0, k_pad = 0;            // Wd is [n_pad, k_pad] int8 row-major (n_pad multiple of 256, k_pad of 128)
  int32_t n128 = 0, n_chunks = 0;          // 128-channel tiles, 128-byte K chunks
  const int8_t* wd = nullptr;
  const uint16_t* klist[2] = {nullptr, nullptr};    // [CG - 1]: live-chunk lists per tile of 128 * CG channels
  const uint16_t* kcount[2] = {nullptr, nullptr};
  int64_t live_chunks[2] = {0, 0};         // sum of the list lengths (tooling)
};

struct accel_plan {
  accel::Plan p;
  std::vector<int32_t> row_ptr, col_idx;   // host copies of the BSR structure (for later re-layouts)
  WsState ws;
  GwState gw;
};

namespace {

bool ws_geometry_ok(const accel_plan* plan, int32_t c_in, int32_t c_out, int32_t ksize) {
  if (ksize == 7)      // stem layout (stem_ws.cuh): K rows (c, kh) fit one 32-row chunk, one lane quadrant slice per image
    return c_in > 0 && c_in * 7 <= accel::kWsCk && c_out > 0 && c_out <= 64 && c_out <= plan->p.nbr * accel::kBlock &&
           (static_cast<int64_t>(c_in) * 49 + accel::kBlock - 1) / accel::kBlock <= plan->p.nbc;
  if ((ksize != 3 && ksize != 1) || c_in <= 0 || c_in % accel::kWsCk != 0 || c_in / accel::kWsCk > accel::kWsMaxChunks) return false;
  if (c_out <= 0 || c_out > plan->p.nbr * accel::kBlock || (c_out + accel::kWsCo - 1) / accel::kWsCo > accel::kWsMaxGroups)
    return false;
  return (static_cast<int64_t>(c_in) * ksize * ksize + accel::kBlock - 1) / accel::kBlock <= plan->p.nbc;
}
size_t ws_blob_bytes(int32_t c_in, int32_t c_out, int32_t taps) {
  if (taps == 49) return accel::kStWBytes;
  return static_cast<size_t>((c_out + accel::kWsCo - 1) / accel::kWsCo) * (c_in / accel::kWsCk) * taps * accel::kWsTapBytes;
}

constexpr int kWsNotApplicable = 1;

// Route a convolution to conv_ws_kernel when geometry, layouts and alignment allow; kWsNotApplicable otherwise.
// 3x3 pad 1, stride 1 or 2.  plan_ds / epi_ds / out_ds (stride 2 only, may be null): the 1x1 / stride 2 / pad 0 convolution of
// the same input (the ResNet downsample), fused as one more tap of the same staged tiles.
int try_conv_ws(const accel_plan* plan, const int8_t* input, const accel_conv_geom* g, const accel_epilogue* epi, void* out,
                const accel_out_layout* lay, cudaStream_t st, const accel_plan* plan_ds = nullptr,
                const accel_epilogue* epi_ds = nullptr, void* out_ds = nullptr) {
  const WsState& W = plan->ws;
  // pointwise: 1x1 / stride 1 / pad 0 (one tap, no halo rows, Z1 only)
  const bool pw = W.ready && W.taps == 1 && g->ksize == 1 && g->stride == 1 && g->pad == 0 && !plan_ds;
  if (g_no_ws || !W.ready || g->batch <= 0) return kWsNotApplicable;
  if (!pw && (W.taps != 9 || g->ksize != 3 || (g->stride != 1 && g->stride != 2) || g->pad != 1)) return kWsNotApplicable;
  if (g->c_in != W.c_in || epi->n_channels != W.c_out) return kWsNotApplicable;
  if (!(epi->flags & ACCEL_OUT_I8) || epi->chan_absmax) return kWsNotApplicable;
  const int stride = g->stride;
  if (stride == 2 && ((g->h | g->w) & 1 || epi->residual)) return kWsNotApplicable;
  // stride 2 with streamed weights (Cin > 128): every tile re-reads the filter from L2, so the tile is made as large as
  // TMEM allows (below).  ACCEL_WS_S2_NARROW keeps the 64-pixel tiles (layer4.0 of ResNet-18: 168 us fused, against
  // 125 us on the gather kernels).
  const bool s2_wide = stride == 2 && W.n_chunks > accel::kWsRing9 && !g_ws_s2_narrow;
  if (plan_ds) {
    const WsState& D = plan_ds->ws;
    if (stride != 2 || !D.ready || D.taps != 1 || D.c_in != W.c_in || D.c_out != W.c_out || !epi_ds || !out_ds) return kWsNotApplicable;
    if (epi_ds->n_channels != W.c_out || !(epi_ds->flags & ACCEL_OUT_I8) || epi_ds->chan_absmax || epi_ds->residual ||
        !epi_ds->chan_scale || (reinterpret_cast<uintptr_t>(out_ds) & 15))
      return kWsNotApplicable;
  }
  const int Wd = g->w, H = g->h;
  const int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? Wd / 2 : Wd;
  const int P = Wd <= 14 ? 16 : (Wd <= 30 ? 32 : (Wd <= 62 ? 64 : 0));     // two padding pixels end every staged row
  if (!P) return kWsNotApplicable;
  const int64_t in_pitch = g->in_row_pitch > 0 ? g->in_row_pitch : g->w;
  if ((in_pitch & 15) || (reinterpret_cast<uintptr_t>(input) & 15)) return kWsNotApplicable;
  // output (and residual): NCHW, 16-byte aligned rows that can take whole 16-pixel stores
  const int w16 = stride == 2 ? (Wo + 7) / 8 * 8 : (Wo + 15) / 16 * 16;     // whole 16-byte (stride 2: 8-byte) stores
  int64_t out_pitch;
  if (lay->row_len == 0) {
    out_pitch = Wo;
    if (Wo % 16 || lay->chan_stride != static_cast<int64_t>(Ho) * Wo) return kWsNotApplicable;
  } else {
    out_pitch = lay->row_pitch;
    if (lay->row_len != Wo || lay->chan_stride != static_cast<int64_t>(Ho) * out_pitch) return kWsNotApplicable;
  }
  if (lay->row_stride != 1 || out_pitch < w16 || (out_pitch & 15) || lay->rows_per_image != static_cast<int64_t>(Ho) * Wo ||
      lay->image_stride != lay->chan_stride * W.c_out)
    return kWsNotApplicable;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(epi->residual) & 15)) return kWsNotApplicable;
  if (lay->image_stride * g->batch >= (1ll << 31) || in_pitch * H * g->c_in * g->batch >= (1ll << 40)) return kWsNotApplicable;
  if (in_pitch * H * accel::kWsCk >= (1ll << 31)) return kWsNotApplicable;

  std::call_once(g_attr_once, set_kernel_attrs);
  if (g_attr_err != cudaSuccess) return cuda_fail(g_attr_err, "cudaFuncSetAttribute(smem)");
  static thread_local accel::WsLaunch L;
  accel::WsParams& p = L.p;
  std::memset(&p, 0, sizeof(p));
  p.C = g->c_in; p.H = H; p.W = Wd; p.B = g->batch; p.P = P;
  int best_r = 1, best_cost = INT_MAX;
  for (int r = 1; r <= (stride == 2 && !s2_wide ? 64 : 128) / P; ++r) {       // stride 2: N <= 64 (three accumulators per set)
    const int cost = (Ho + r - 1) / r * r;
    const int copies = accel::kWsCk * (stride == 2 ? 2 * r + 1 : (pw ? r : r + 2)) * (P >> 4);      // 16-byte loader copies per stage
    if (copies > accel::kWsLoadThreads * accel::kWsLoadOps) break;
    if (cost <= best_cost) { best_cost = cost; best_r = r; }
  }
  p.R = best_r; p.N = p.R * P;
  p.stride = stride; p.Ho = Ho; p.Wo = Wo;

  p.rows_in = stride == 2 ? 2 * p.R + 1 : (pw ? p.R : p.R + 2);
  p.pw = pw ? 1 : 0; p.ypad = pw ? 0 : 1;
  p.w_src_stride = pw ? accel::kWsTapBytes : accel::kWsChunkBytes;
  p.has_ds = plan_ds ? 1 : 0;
  p.acc_single = (s2_wide && plan_ds) ? 1 : 0;
  p.v_col = p.acc_single ? 256 : 128;
  p.w_chunk_bytes = pw ? accel::kWsTapBytes : accel::kWsChunkBytes + (plan_ds ? accel::kWsTapBytes : 0);
  p.n_chunks = W.n_chunks; p.n_groups = W.n_groups; p.c_out = W.c_out;
  p.tiles_per_image = (Ho + p.R - 1) / p.R;
  // 16-pixel rows with <= 7 valid pixels (layer4 of ResNet-18): two images side by side in every staged row
  p.twin = (stride == 1 && P == 16 && Wd <= 7 && W.c_out > 64 && g->batch >= 2 && !g_no_twin &&
            static_cast<int64_t>(g->c_in) * H * in_pitch < (1ll << 30)) ? 1 : 0;
  p.u_alias = p.twin ? 0 : 1;      // twin tiles end in a valid pixel of the second image: U cannot alias Z1's tail (N <= 112)
  p.u_off = p.u_alias ? p.N - 2 : p.N;
  const int64_t n_tiles = static_cast<int64_t>(p.twin ? (g->batch + 1) / 2 : g->batch) * p.tiles_per_image;
  if (n_tiles > INT_MAX) return kWsNotApplicable;
  p.n_tiles = static_cast<int32_t>(n_tiles);
  // 3x3: 36 KB per chunk - resident up to Cin 128, else a ring of four slots; 1x1: 4 KB per chunk - resident up to Cin 1024
  const int ring = pw ? 16 : accel::kWsRing9, resident_max = pw ? accel::kWsMaxWSlots : accel::kWsRing9;
  p.w_resident = W.n_chunks <= resident_max ? 1 : 0;
  p.w_slots = p.w_resident ? W.n_chunks : ring;
  p.a_box_bytes = accel::kWsCk * p.rows_in * P;
  p.a_stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
  const int fixed = 1024 /* alignment slack */ + accel::kWsSmemBar + p.w_slots * p.w_chunk_bytes;
  int a_slots = (kSmemWs - fixed) / p.a_stage_bytes;
  if (a_slots > accel::kWsMaxASlots) a_slots = accel::kWsMaxASlots;
  if (a_slots < 3) return kWsNotApplicable;
  p.a_slots = a_slots;
  p.row_stride = static_cast<uint32_t>(accel::kWsCk * P);          // bytes between staged image rows
  const uint32_t grp_stride = static_cast<uint32_t>(8 * P);        // bytes between 8-channel groups
  p.b_layout = P == 64 ? 4u : (P == 32 ? 6u : 0u);
  // swizzled: LBO = stride between pixel atoms (image rows), SBO = stride between 8-channel groups; unswizzled: swapped
  p.b_lbo = p.b_layout ? p.row_stride * stride : grp_stride;     // stride 2: the window takes every second staged row
  p.b_sbo = p.b_layout ? grp_stride : p.row_stride * stride;
  p.d_tpi = accel::make_fastdiv(static_cast<uint32_t>(p.tiles_per_image));
  p.wblob = W.blob;
  p.x = input; p.in_pitch = static_cast<int32_t>(in_pitch);
  p.epi = *epi;
  if (epi->residual) {
    p.res_fast = residual_divide_mode(epi->res_scale_main, epi->res_scale_res, epi->res_scale_out);
    p.res_rcp = 1.0f / epi->res_scale_out;
  }
  p.out = static_cast<int8_t*>(out);
  p.out_pitch = static_cast<int32_t>(out_pitch);
  p.chan_stride = static_cast<int32_t>(lay->chan_stride);
  p.x_store_end = w16;
  p.image_stride = lay->image_stride;
  std::memcpy(p.masks, W.masks, sizeof(p.masks));
  // Activation stages by TMA tensor tiles (everything but the twin tiles, whose 8-byte half rows are below TMA's 16-byte box
  // minimum).  Same-box A/B, whole ResNet-18 at batch 256: LDGSTS loaders 261.3 k img/s, TMA for 64-byte rows 268.7 k (layer1
  // 64 -> 56 us, layer2.0 + downsample 81 -> 66 us), also for 32-byte rows 271.5 k (layer2 45 -> 42 us), also for 16-byte rows
  // 272.8 k.  tools/ws_timeline.py had shown the two loader warps latency-bound on their own instruction stream (~100
  // instructions per 8 KB stage, 790 cycles alone and ~1100 next to the epilogue warps); the TMA unit needs ~4 cycles per box row
  // whatever the SM's warps are doing.
  p.use_tma = 0;
  if (!p.twin && ((g_ws_tma >= 1 && P == 64) || (g_ws_tma >= 2 && P == 32) || (g_ws_tma >= 3 && P == 16))) {
    const uint64_t dims[4] = {static_cast<uint64_t>(Wd), static_cast<uint64_t>(g->c_in), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(g->batch)};
    const uint64_t strides[3] = {static_cast<uint64_t>(in_pitch) * H, static_cast<uint64_t>(in_pitch),
                                 static_cast<uint64_t>(in_pitch) * H * g->c_in};
    const uint32_t box[4] = {static_cast<uint32_t>(P), static_cast<uint32_t>(accel::kWsCk), static_cast<uint32_t>(p.rows_in), 1u};
    const CUtensorMapSwizzle sw = P == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (P == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    if (p.rows_in <= 256 && encode_tmap(&L.tmap, input, 4, dims, strides, box, sw)) p.use_tma = 1;
  }
  p.dbg = g_dbg_flags;
  p.timeline = g_timeline;
  p.dual = (W.c_out <= 64 && stride == 1) ? 1 : 0;
  if (plan_ds) {
    p.wblob2 = plan_ds->ws.blob;
    p.epi2 = *epi_ds;
    p.out2 = static_cast<int8_t*>(out_ds);
    if (plan_ds->ws.n_chunks > accel::kWsMaxChunks2 || plan_ds->ws.n_groups > accel::kWsMaxGroups2) return kWsNotApplicable;
    for (int gi = 0; gi < plan_ds->ws.n_groups; ++gi)
      for (int j = 0; j < plan_ds->ws.n_chunks; ++j)
        p.masks2[gi * accel::kWsMaxChunks2 + j] = plan_ds->ws.masks[gi * accel::kWsMaxChunks + j];
  }
  const int n_items = p.dual ? (p.n_tiles + 1) / 2 : p.n_tiles;
  int per_group = sm_count() / p.n_groups;
  if (per_group > n_items) per_group = n_items;
  if (per_group < 1) return kWsNotApplicable;
  const int smem = fixed + p.a_slots * p.a_stage_bytes;
  const int resmode = !epi->residual ? 0 : (p.res_fast == 3 ? 4 : (p.res_fast == 2 ? 3 : (p.res_fast == 1 ? 1 : 2)));
  const int mode = stride == 2 ? accel::kWsModeS2
                               : (p.twin ? (p.pw ? accel::kWsModeTwinPw : accel::kWsModeTwin) : (p.pw ? accel::kWsModePw : accel::kWsModeS1));
  // conversion-free epilogue: the caller promised |accumulator + bias| < 2^22 (both launches of a fused stride-2 pair must)
  const bool fast = !g_no_fast_epi && epi->acc_bound > 0 && epi->acc_bound < (1 << 22) && resmode == 4 && !plan_ds;
  WsKernelFn kfn = ws_kernel_ptr(mode, resmode, epi->sat_count != nullptr, fast);
  if (!kfn) return kWsNotApplicable;
  cudaError_t e = launch_overlapped(kfn, static_cast<unsigned>(per_group * p.n_groups), accel::kWsThreads, smem, st, L);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv_ws_kernel launch");
  ++g_ws_launches;
  g_ws_fast_launches += fast ? 1 : 0;
  return ACCEL_OK;
}


// Route a GEMM to gemm_ws_kernel when the plan carries the dense-equivalent layout and the tensors allow TMA
// (16-byte aligned activation rows, a strided output layout without padded image rows); kWsNotApplicable otherwise.
int try_gemm_ws(const accel_plan* plan, const int8_t* act, int64_t M, int64_t K, int64_t lda, const accel_epilogue* epi, void* out,
                const accel_out_layout* lay, cudaStream_t st) {
  const GwState& G = plan->gw;
  if (g_no_gemm_ws || !G.ready || M <= 0 || K <= 0) return kWsNotApplicable;
  if ((reinterpret_cast<uintptr_t>(act) & 15) || (lda & 15) || lay->row_len != 0 || epi->chan_absmax) return kWsNotApplicable;
  if (epi->n_channels <= 0) return kWsNotApplicable;
  const int cg = g_gemm_ws_cg == 1 ? 1 : 2;
  std::call_once(g_attr_once, set_kernel_attrs);
  if (g_attr_err != cudaSuccess) return cuda_fail(g_attr_err, "cudaFuncSetAttribute(smem)");
  static thread_local accel::GwLaunch L;
  accel::GwParams& p = L.p;
  std::memset(&p, 0, sizeof(p));
  const uint64_t wdims[2] = {static_cast<uint64_t>(G.k_pad), static_cast<uint64_t>(G.n_pad)};
  const uint64_t wstr[1] = {static_cast<uint64_t>(G.k_pad)};
  const uint64_t xdims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  const uint64_t xstr[1] = {static_cast<uint64_t>(lda)};
  const uint32_t box[2] = {static_cast<uint32_t>(accel::kGwKc), 128u};
  if (!encode_tmap(&L.tmap_w, G.wd, 2, wdims, wstr, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
      !encode_tmap(&L.tmap_x, act, 2, xdims, xstr, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
    return kWsNotApplicable;
  p.M = M;
  // channel tiles: only those that hold channels the caller wants written
  const int tile_ch = accel::kGwCh * cg;
  int n_ct = (epi->n_channels + tile_ch - 1) / tile_ch;
  const int n_ct_plan = cg == 2 ? G.n128 / 2 : G.n128;
  if (n_ct > n_ct_plan) return kWsNotApplicable;       // n_channels <= nbr * 14 <= n_pad: cannot happen
  p.n_ch_tiles = n_ct;
  const int64_t n_rt = (M + accel::kGwRows - 1) / accel::kGwRows;
  if (n_rt * n_ct >= (1ll << 31)) return kWsNotApplicable;
  p.n_row_tiles = static_cast<int32_t>(n_rt);
  const int stage_bytes = accel::kGwBoxBytes * (cg == 2 ? 2 : 3);
  int stages = (kSmemWs - 1024 - accel::kGwSmemBar) / stage_bytes;
  if (stages > accel::kGwMaxStages) stages = accel::kGwMaxStages;
  if (g_gemm_ws_stages >= 2 && g_gemm_ws_stages < stages) stages = g_gemm_ws_stages;
  p.n_stages = stages;
  p.klist = G.klist[cg - 1]; p.kcount = G.kcount[cg - 1]; p.klist_stride = G.n_chunks;
  p.k_chunks_x = static_cast<int32_t>((K + accel::kGwKc - 1) / accel::kGwKc);
  p.epi = *epi;
  if (epi->residual) {
    p.res_fast = residual_fast_divide_ok(epi->res_scale_main, epi->res_scale_res, epi->res_scale_out) ? 1 : 0;
    p.res_rcp = 1.0f / epi->res_scale_out;
  }
  p.out = out; p.lay = *lay;
  p.single_image = lay->rows_per_image >= M ? 1 : 0;
  p.d_rpi = accel::make_fastdiv(static_cast<uint32_t>(lay->rows_per_image < (1ll << 31) ? lay->rows_per_image : 1));
  p.dbg = g_dbg_flags;
  const int outk = (epi->flags & ACCEL_OUT_I32) ? 0 : ((epi->flags & ACCEL_OUT_I8) ? 1 : 2);
  const int64_t tiles = n_rt * n_ct;
  const int units_max = cg == 2 ? sm_count() / 2 : sm_count();
  const int units = static_cast<int>(tiles < units_max ? tiles : units_max);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(units * cg)); cfg.blockDim = dim3(accel::kGwThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(1024 + accel::kGwSmemBar + stages * stage_bytes); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cg); attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gw_kernel_ptr(cg, outk), L);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_ws_kernel launch");
  ++g_gw_launches;
  return ACCEL_OK;
}

}  // namespace

extern "C" {

const char* accel_last_error_string(void) { return g_err.c_str(); }
const char* accel_version(void) { return "accel_b200 0.1 (sm_100a, tcgen05 kind::i8)"; }

void accel_debug_set_timeline(long long* dev_buffer) {
  g_timeline = dev_buffer;
  const char* f = std::getenv("ACCEL_DBG_FLAGS");
  g_dbg_flags = f ? std::atoi(f) : 0;
  g_no_fast_epi = std::getenv("ACCEL_NO_FAST_EPI") != nullptr;
  g_ws_tma = std::getenv("ACCEL_WS_TMA") ? std::atoi(std::getenv("ACCEL_WS_TMA")) : 3;
  g_stem_rows = std::getenv("ACCEL_STEM_ROWS") ? std::atoi(std::getenv("ACCEL_STEM_ROWS")) : 0;
}

long long accel_debug_counter(int which) {
  return which == 0 ? g_ws_launches : which == 1 ? g_gw_launches : which == 2 ? g_ws_fast_launches : -1;
}

int accel_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return fail(ACCEL_INIT_FAILED, "no CUDA device");
  if (prop.major != 10) return fail(ACCEL_INIT_FAILED, "device is not sm_100 (tcgen05 kind::i8 needs a B200)");
  return ACCEL_OK;
}

int accel_plan_create(const int32_t* row_ptr_host, const int32_t* col_idx_host, int32_t n_block_rows,
                      int32_t n_block_cols, int32_t block, int32_t group_rows_hint, accel_plan** plan_out,
                      size_t* workspace_bytes) {
  if (!row_ptr_host || !plan_out || !workspace_bytes) return fail(ACCEL_INVALID_CONFIG, "null argument");
  if (block != accel::kBlock) return fail(ACCEL_INVALID_CONFIG, "Block size must be 14");  // accel.py:196
  if (n_block_rows < 0 || n_block_cols < 0) return fail(ACCEL_INVALID_CONFIG, "negative block grid");
  if (row_ptr_host[n_block_rows] > 0 && !col_idx_host) return fail(ACCEL_INVALID_CONFIG, "null col_idx");
  accel_plan* pl = new accel_plan();
  pl->p.group_rows = group_rows_hint;
  const std::string msg = accel::build_plan(row_ptr_host, col_idx_host, n_block_rows, n_block_cols, &pl->p);
  if (!msg.empty()) {
    delete pl;
    return fail(ACCEL_INVALID_CONFIG, msg);
  }
  pl->row_ptr.assign(row_ptr_host, row_ptr_host + n_block_rows + 1);
  if (row_ptr_host[n_block_rows] > 0) pl->col_idx.assign(col_idx_host, col_idx_host + row_ptr_host[n_block_rows]);
  *plan_out = pl;
  *workspace_bytes = pl->p.ws_bytes;
  return ACCEL_OK;
}

int accel_plan_conv_ws_bytes(const accel_plan* plan, int32_t c_in, int32_t c_out, int32_t ksize, size_t* bytes) {
  if (!plan || !bytes) return fail(ACCEL_INVALID_CONFIG, "null argument");
  *bytes = 0;
  if (!ws_geometry_ok(plan, c_in, c_out, ksize)) return ACCEL_OK;          // 0 bytes: this geometry has no such path
  *bytes = ws_blob_bytes(c_in, c_out, ksize * ksize) + (static_cast<size_t>(plan->p.nnz) * 8 + 255) / 256 * 256;
  return ACCEL_OK;
}

void accel_plan_conv_ws_release(accel_plan* plan) {
  if (plan) { plan->ws.ready = false; plan->ws.blob = nullptr; }
}

int accel_plan_conv_ws_prepare(accel_plan* plan, const int8_t* blocks_dev, int32_t c_in, int32_t c_out, int32_t ksize,
                               void* workspace_dev, size_t workspace_bytes, accel_stream_t stream) {
  if (!plan || !workspace_dev) return fail(ACCEL_INVALID_CONFIG, "null plan / workspace");
  if (!ws_geometry_ok(plan, c_in, c_out, ksize)) return fail(ACCEL_INVALID_CONFIG, "geometry has no weight-stationary path");
  const int32_t taps = ksize * ksize;
  const size_t blob = ws_blob_bytes(c_in, c_out, taps);
  const int64_t nnz = plan->p.nnz;
  if (workspace_bytes < blob + static_cast<size_t>(nnz) * 8) return fail(ACCEL_MEMORY_ERROR, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 1023) return fail(ACCEL_MEMORY_ERROR, "workspace not 1024-byte aligned");
  if (nnz > 0 && !blocks_dev) return fail(ACCEL_INVALID_CONFIG, "null blocks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WsState& W = plan->ws;
  W.ready = false;
  W.c_in = c_in; W.c_out = c_out; W.taps = taps; W.n_chunks = taps == 49 ? 1 : c_in / accel::kWsCk; W.n_groups = (c_out + accel::kWsCo - 1) / accel::kWsCo;
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  CU(cudaMemsetAsync(ws, 0, blob, st));
  std::memset(W.masks, 0, sizeof(W.masks));
  std::vector<int32_t> blk_row(static_cast<size_t>(nnz));
  const int32_t K = c_in * taps;
  for (int32_t br = 0; br < plan->p.nbr; ++br)
    for (int32_t i = plan->row_ptr[br]; i < plan->row_ptr[br + 1]; ++i) {
      blk_row[i] = br;
      if (taps == 49) continue;
      const int32_t bc = plan->col_idx[i];
      const int co_lo = br * accel::kBlock, co_hi = std::min(co_lo + accel::kBlock, c_out) - 1;
      if (co_hi < co_lo) continue;
      for (int k = bc * accel::kBlock; k < std::min((bc + 1) * accel::kBlock, K); ++k) {
        const int c = k / taps, tap = k % taps, j = c / accel::kWsCk;
        W.masks[(co_lo / accel::kWsCo) * accel::kWsMaxChunks + j] |= static_cast<uint16_t>(1u << tap);
        W.masks[(co_hi / accel::kWsCo) * accel::kWsMaxChunks + j] |= static_cast<uint16_t>(1u << tap);
      }
    }
  // the first chunk always issues all nine taps: they initialise the accumulators in a fixed order
  for (int gi = 0; gi < W.n_groups; ++gi) W.masks[gi * accel::kWsMaxChunks] = static_cast<uint16_t>((1u << taps) - 1u);
  if (nnz > 0) {
    int32_t* d_row = reinterpret_cast<int32_t*>(ws + blob);
    int32_t* d_col = d_row + nnz;
    CU(cudaMemcpyAsync(d_row, blk_row.data(), static_cast<size_t>(nnz) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_col, plan->col_idx.data(), static_cast<size_t>(nnz) * 4, cudaMemcpyHostToDevice, st));
    if (taps == 49)
      accel::stem_scatter_kernel<<<static_cast<unsigned>(nnz), 64, 0, st>>>(blocks_dev, d_row, d_col, nnz, c_in, c_out, ws);
    else
      accel::ws_scatter_kernel<<<static_cast<unsigned>(nnz), 64, 0, st>>>(blocks_dev, d_row, d_col, nnz, c_in, c_out, W.n_chunks, taps, ws);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(st));     // blk_row is pageable host memory
  W.blob = ws;
  W.ready = true;
  return ACCEL_OK;
}

static size_t gw_align(size_t v) { return (v + 255) / 256 * 256; }

int accel_plan_gemm_ws_bytes(const accel_plan* plan, size_t* bytes) {
  if (!plan || !bytes) return fail(ACCEL_INVALID_CONFIG, "null argument");
  const int64_t n_pad = (static_cast<int64_t>(plan->p.nbr) * accel::kBlock + 255) / 256 * 256;
  const int64_t k_pad = (static_cast<int64_t>(plan->p.nbc) * accel::kBlock + 127) / 128 * 128;
  *bytes = 0;
  if (n_pad == 0 || k_pad == 0 || k_pad / 128 > 65535) return ACCEL_OK;      // nothing to multiply / chunk index overflows u16
  const size_t n128 = static_cast<size_t>(n_pad / 128), nch = static_cast<size_t>(k_pad / 128);
  *bytes = gw_align(static_cast<size_t>(n_pad) * k_pad) + gw_align(n128 * nch) /* flags */ +
           gw_align(n128 * nch * 2) + gw_align(n128 * 2) + gw_align(n128 / 2 * nch * 2) + gw_align(n128) +
           gw_align(static_cast<size_t>(plan->p.nnz) * 8);
  return ACCEL_OK;
}

int accel_plan_gemm_ws_prepare(accel_plan* plan, const int8_t* blocks_dev, void* workspace_dev, size_t workspace_bytes,
                               accel_stream_t stream) {
  if (!plan || !workspace_dev) return fail(ACCEL_INVALID_CONFIG, "null plan / workspace");
  size_t need = 0;
  accel_plan_gemm_ws_bytes(plan, &need);
  if (need == 0) return fail(ACCEL_INVALID_CONFIG, "plan has no dense-equivalent GEMM layout");
  if (workspace_bytes < need) return fail(ACCEL_MEMORY_ERROR, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 255) return fail(ACCEL_MEMORY_ERROR, "workspace not 256-byte aligned");
  const int64_t nnz = plan->p.nnz;
  if (nnz > 0 && !blocks_dev) return fail(ACCEL_INVALID_CONFIG, "null blocks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GwState& G = plan->gw;
  G.ready = false;
  G.n_pad = (static_cast<int64_t>(plan->p.nbr) * accel::kBlock + 255) / 256 * 256;
  G.k_pad = (static_cast<int64_t>(plan->p.nbc) * accel::kBlock + 127) / 128 * 128;
  G.n128 = static_cast<int32_t>(G.n_pad / 128); G.n_chunks = static_cast<int32_t>(G.k_pad / 128);
  const size_t n128 = static_cast<size_t>(G.n128), nch = static_cast<size_t>(G.n_chunks);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  int8_t* wd = reinterpret_cast<int8_t*>(ws);
  uint8_t* flags = ws + gw_align(static_cast<size_t>(G.n_pad) * G.k_pad);
  uint16_t* kl1 = reinterpret_cast<uint16_t*>(flags + gw_align(n128 * nch));
  uint16_t* kc1 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(kl1) + gw_align(n128 * nch * 2));
  uint16_t* kl2 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(kc1) + gw_align(n128 * 2));
  uint16_t* kc2 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(kl2) + gw_align(n128 / 2 * nch * 2));
  int32_t* d_row = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(kc2) + gw_align(n128));
  CU(cudaMemsetAsync(wd, 0, static_cast<size_t>(G.n_pad) * G.k_pad, st));
  std::vector<int32_t> blk_row(static_cast<size_t>(nnz));
  for (int32_t br = 0; br < plan->p.nbr; ++br)
    for (int32_t i = plan->row_ptr[br]; i < plan->row_ptr[br + 1]; ++i) blk_row[i] = br;
  if (nnz > 0) {
    int32_t* d_col = d_row + nnz;
    CU(cudaMemcpyAsync(d_row, blk_row.data(), static_cast<size_t>(nnz) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_col, plan->col_idx.data(), static_cast<size_t>(nnz) * 4, cudaMemcpyHostToDevice, st));
    accel::gw_scatter_kernel<<<static_cast<unsigned>(nnz), 64, 0, st>>>(blocks_dev, d_row, d_col, nnz, G.k_pad, wd);
    CU(cudaGetLastError());
  }
  accel::gw_flags_kernel<<<dim3(static_cast<unsigned>(nch), static_cast<unsigned>(n128)), 128, 0, st>>>(wd, G.k_pad, G.n_chunks, flags);
  CU(cudaGetLastError());
  std::vector<uint8_t> hflags(n128 * nch);
  CU(cudaMemcpyAsync(hflags.data(), flags, n128 * nch, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  // live-chunk lists: per 128-channel tile (CG = 1) and per 256-channel tile (CG = 2)
  std::vector<uint16_t> l1(n128 * nch, 0), c1(n128, 0), l2(n128 / 2 * nch, 0), c2(n128 / 2, 0);
  G.live_chunks[0] = G.live_chunks[1] = 0;
  for (size_t t = 0; t < n128; ++t)
    for (size_t j = 0; j < nch; ++j)
      if (hflags[t * nch + j]) l1[t * nch + c1[t]++] = static_cast<uint16_t>(j);
  for (size_t t = 0; t < n128 / 2; ++t)
    for (size_t j = 0; j < nch; ++j)
      if (hflags[2 * t * nch + j] | hflags[(2 * t + 1) * nch + j]) l2[t * nch + c2[t]++] = static_cast<uint16_t>(j);
  for (size_t t = 0; t < n128; ++t) G.live_chunks[0] += c1[t];
  for (size_t t = 0; t < n128 / 2; ++t) G.live_chunks[1] += c2[t];
  CU(cudaMemcpyAsync(kl1, l1.data(), l1.size() * 2, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(kc1, c1.data(), c1.size() * 2, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(kl2, l2.data(), l2.size() * 2, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(kc2, c2.data(), c2.size() * 2, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));     // the host vectors are pageable
  G.wd = wd;
  G.klist[0] = kl1; G.kcount[0] = kc1; G.klist[1] = kl2; G.kcount[1] = kc2;
  G.ready = true;
  return ACCEL_OK;
}

void accel_plan_gemm_ws_release(accel_plan* plan) {
  if (plan) plan->gw.ready = false;
}

int64_t accel_plan_gemm_ws_live_chunks(const accel_plan* plan, int32_t cta_group) {
  if (!plan || !plan->gw.ready || cta_group < 1 || cta_group > 2) return -1;
  return plan->gw.live_chunks[cta_group - 1];
}

int accel_plan_upload(accel_plan* plan, const int8_t* blocks_dev, void* workspace_dev, size_t workspace_bytes,
                      accel_stream_t stream) {
  if (!plan || !workspace_dev) return fail(ACCEL_INVALID_CONFIG, "null plan / workspace");
  accel::Plan& P = plan->p;
  if (workspace_bytes < P.ws_bytes) return fail(ACCEL_MEMORY_ERROR, "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 255) return fail(ACCEL_MEMORY_ERROR, "workspace not 256-byte aligned");
  if (P.nnz > 0 && !blocks_dev) return fail(ACCEL_INVALID_CONFIG, "null blocks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  if (P.n_tiles > 0) {
    CU(cudaMemcpyAsync(ws + P.off_tilesrc, P.tile_src.data(), P.tile_src.size() * sizeof(accel::TileSrc),
                       cudaMemcpyHostToDevice, st));
    accel::repack_blocks_kernel<<<static_cast<unsigned>(P.n_tiles), 128, 0, st>>>(
        blocks_dev, ws + P.off_blob, reinterpret_cast<const accel::TileSrc*>(ws + P.off_tilesrc), P.n_tiles);
    CU(cudaGetLastError());
  }
  // the host vector is pageable: make sure the copy has consumed it before returning
  CU(cudaStreamSynchronize(st));
  P.ws_dev = ws;
  P.uploaded = true;
  return ACCEL_OK;
}

void accel_plan_destroy(accel_plan* plan) { delete plan; }
int64_t accel_plan_num_blocks(const accel_plan* plan) { return plan ? plan->p.nnz : 0; }
int64_t accel_plan_num_mma(const accel_plan* plan) { return plan ? static_cast<int64_t>(plan->p.ops.size()) : 0; }
int64_t accel_plan_num_tiles(const accel_plan* plan) { return plan ? plan->p.n_tiles : 0; }

int64_t accel_plan_export_ops(const accel_plan* plan, int32_t* rec, int64_t cap) {
  if (!plan) return 0;
  const accel::Plan& P = plan->p;
  int64_t t = 0;
  for (size_t gi = 0; gi < P.groups.size(); ++gi) {
    const accel::GroupRec& G = P.groups[gi];
    for (uint32_t b = G.batch_begin; b < G.batch_end; ++b)
      for (uint32_t i = 0; i < accel::batch_tiles(P.batches[b]); ++i, ++t) {
        if (rec && t < cap) {
          int32_t* r = rec + t * 8;
          r[0] = static_cast<int32_t>(gi); r[1] = static_cast<int32_t>(G.br0_rows & 0xffffu); r[2] = P.tile_dbg[t * 4 + 1];
          r[3] = P.tile_dbg[t * 4 + 2]; r[4] = P.tile_dbg[t * 4 + 3]; r[5] = P.tile_src[t].blk_lo;
          r[6] = P.tile_src[t].blk_hi; r[7] = static_cast<int32_t>(b);
        }
      }
  }
  return t;
}

int64_t accel_plan_export_mma(const accel_plan* plan, int32_t* rec, int64_t cap) {
  if (!plan) return 0;
  const accel::Plan& P = plan->p;
  int64_t n = 0, tile0 = 0;
  for (size_t gi = 0; gi < P.groups.size(); ++gi) {
    const accel::GroupRec& G = P.groups[gi];
    uint32_t op = G.op_begin;
    for (uint32_t b = G.batch_begin; b < G.batch_end; ++b) {
      for (uint32_t i = 0; i < accel::batch_runs(P.batches[b]); ++i, ++op, ++n) {
        if (rec && n < cap) {
          int32_t* r = rec + n * 8;
          const accel::OpRec& o = P.ops[op];
          r[0] = static_cast<int32_t>(gi); r[1] = static_cast<int32_t>(b);
          r[2] = static_cast<int32_t>(o.d_n & 0x1ffu);                 // accumulator column
          r[3] = static_cast<int32_t>(((o.d_n >> 17) & 0x3fu) * 8);    // N
          r[4] = static_cast<int32_t>(o.a_b & 0x1ffu);                 // activation column inside the stage
          r[5] = static_cast<int32_t>(tile0 + (o.a_b >> 16) / (accel::kBTileBytes / 16));   // first B tile (blob order)
          r[6] = static_cast<int32_t>(accel::batch_chunk(P.batches[b]));
          r[7] = static_cast<int32_t>(G.br0_rows & 0xffffu);
        }
      }
      tile0 += accel::batch_tiles(P.batches[b]);
    }
  }
  return n;
}

int accel_bsr_gemm_i8(const accel_plan* plan, const int8_t* act, int64_t M, int64_t K, int64_t lda,
                      const accel_epilogue* epi, void* out, const accel_out_layout* layout, accel_stream_t stream) {
  if (!plan) return fail(ACCEL_INVALID_CONFIG, "null plan");
  if (!plan->p.uploaded) return fail(ACCEL_NOT_READY, "Weights not loaded");  // accel.py:296
  if (M < 0 || K < 0 || lda < K) return fail(ACCEL_INVALID_CONFIG, "bad activation shape");
  if (M > INT_MAX) return fail(ACCEL_INVALID_CONFIG, "more than 2^31 activation rows");
  if (M > 0 && !act) return fail(ACCEL_INVALID_CONFIG, "Activations not loaded");  // accel.py:297
  if (K > INT_MAX / 2) return fail(ACCEL_INVALID_CONFIG, "K too large");
  // accumulator overflow guard (SURVEY.md hard part 7): |acc| <= K*128*128 must stay below 2^31
  if (K >= 131072) return fail(ACCEL_INVALID_CONFIG, "K >= 131072 could overflow the INT32 accumulator");
  int rc = check_epilogue(epi, layout, out, plan->p.nbr * accel::kBlock);
  if (rc) return rc;
  rc = try_gemm_ws(plan, act, M, K, lda, epi, out, layout, static_cast<cudaStream_t>(stream));
  if (rc != kWsNotApplicable) return rc;
  accel::TcParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.x = act; prm.M = M; prm.K = static_cast<int32_t>(K); prm.lda = lda;
  prm.x_align2 = ((reinterpret_cast<uintptr_t>(act) | static_cast<uintptr_t>(lda)) & 1) == 0;
  prm.epi = *epi; prm.out = out; prm.lay = *layout;
  const bool ring = ((reinterpret_cast<uintptr_t>(act) | static_cast<uintptr_t>(lda)) & 15) == 0;
  if (ring) {   // 16-byte aligned rows: streamed into the shared-memory ring, by TMA when a descriptor can be made
    prm.ring_slots = accel::kMaxRingSlots;
    prm.slot_bytes = accel::kGemmSlotBytes;
    CUtensorMap tm;
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(lda)};
    const uint32_t box[2] = {144, accel::kTileM};
    prm.use_tma = (M > 0 && K > 0 && encode_tmap(&tm, act, 2, dims, strides, box)) ? 1 : 0;
    return launch_tc(&plan->p, prm, accel::kModeGemm, accel::kSmemRing + prm.ring_slots * prm.slot_bytes + accel::kRingSlack,
                     static_cast<cudaStream_t>(stream), prm.use_tma ? &tm : nullptr);
  }
  return launch_tc(&plan->p, prm, accel::kModeDirect, accel::kSmemRing, static_cast<cudaStream_t>(stream));
}

int accel_conv_bsr_i8(const accel_plan* plan, const int8_t* input_nchw, const accel_conv_geom* g,
                      const accel_epilogue* epi, void* out, const accel_out_layout* layout, accel_stream_t stream) {
  if (!plan || !g) return fail(ACCEL_INVALID_CONFIG, "null plan / geometry");
  if (!plan->p.uploaded) return fail(ACCEL_NOT_READY, "Weights not loaded");
  if (g->batch < 0 || g->c_in <= 0 || g->h <= 0 || g->w <= 0 || g->ksize <= 0 || g->stride <= 0 || g->pad < 0)
    return fail(ACCEL_INVALID_CONFIG, "bad convolution geometry");
  if (g->h + 2 * g->pad < g->ksize || g->w + 2 * g->pad < g->ksize)
    return fail(ACCEL_INVALID_CONFIG, "kernel larger than padded input");
  const int64_t K = static_cast<int64_t>(g->c_in) * g->ksize * g->ksize;
  if (K >= 131072) return fail(ACCEL_INVALID_CONFIG, "K >= 131072 could overflow the INT32 accumulator");
  if ((K + accel::kBlock - 1) / accel::kBlock > plan->p.nbc)
    return fail(ACCEL_INVALID_CONFIG, "Cin*k*k does not match the plan's K tiles");
  if (g->batch > 0 && !input_nchw) return fail(ACCEL_INVALID_CONFIG, "Activations not loaded");
  int rc = check_epilogue(epi, layout, out, plan->p.nbr * accel::kBlock);
  if (rc) return rc;
  rc = try_conv_ws(plan, input_nchw, g, epi, out, layout, static_cast<cudaStream_t>(stream));
  if (rc != kWsNotApplicable) return rc;
  accel::TcParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.Ho = (g->h + 2 * g->pad - g->ksize) / g->stride + 1;
  prm.Wo = (g->w + 2 * g->pad - g->ksize) / g->stride + 1;
  prm.x = input_nchw; prm.M = static_cast<int64_t>(g->batch) * prm.Ho * prm.Wo; prm.K = static_cast<int32_t>(K);
  prm.C = g->c_in; prm.H = g->h; prm.W = g->w; prm.ksz = g->ksize; prm.stride = g->stride; prm.pad = g->pad;
  prm.Wp = g->in_row_pitch > 0 ? g->in_row_pitch : g->w;
  if (prm.Wp < g->w) return fail(ACCEL_INVALID_CONFIG, "in_row_pitch smaller than the row");
  if (prm.M > INT_MAX) return fail(ACCEL_INVALID_CONFIG, "more than 2^31 output positions");
  prm.conv = 1;
  prm.x_small = static_cast<int64_t>(g->batch) * g->c_in * g->h * prm.Wp < (1ll << 32);
  prm.epi = *epi; prm.out = out; prm.lay = *layout;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((g->ksize == 3 || g->ksize == 7) && g->pad <= 3) {
    // Shared-memory ring of the input rows one stage needs: [channel plane][input row][pitch], each row once.
    const int ks = g->ksize, gps = 126 / ks;
    const int64_t out_rows_total = static_cast<int64_t>(g->batch) * prm.Ho;
    int ro = (accel::kTileM - 1) / prm.Wo + 2;                       // output rows one 128-row tile can touch
    if (ro > out_rows_total) ro = static_cast<int>(out_rows_total);
    int ni = (ro - 1) / prm.Ho + 2;                                  // images it can touch
    if (ni > ro) ni = ro;
    const int extra = ks > g->stride ? ks - g->stride : 0;
    const int hr = (ro - 1) * g->stride + (ni - 1) * extra + ks;      // input rows per channel plane (upper bound)
    // copy granularity: 16-byte cp.async when every input row starts 16-byte aligned, 4-byte cp.async when rows
    // are 4-byte aligned, else 32-bit loads through registers
    const uintptr_t xa = reinterpret_cast<uintptr_t>(input_nchw);
    const int vec = ((xa | static_cast<uintptr_t>(prm.Wp)) & 15) == 0 ? 16
                    : (((xa | static_cast<uintptr_t>(prm.Wp)) & 3) == 0 && g->w % 4 == 0) ? 4 : 0;
    int pitch, lpad, ipr;
    if (vec == 16) {
      const int w16 = ((g->w + 15) / 16) * 16;
      lpad = 16; ipr = w16 / 16;
      pitch = 16 + w16 + ((w16 - g->w >= g->pad) ? 0 : 16);
    } else {
      lpad = 4; ipr = (g->w + 3) / 4;
      pitch = ((g->w + 3) / 4) * 4 + 8;                              // 4 zero bytes left, >= 4 right (pad <= 3)
    }
    const int nch = (gps % ks == 0) ? gps / ks : (gps + ks - 2) / ks + 1;
    const int fixed = accel::kSmemRing + accel::kRingSlack;
    auto pick_slots = [&](int64_t slot_bytes) {
      if (fixed + 3 * slot_bytes <= kSmemTwoCtas) return 3;
      if (fixed + 2 * slot_bytes <= kSmemTwoCtas) return 2;
      if (fixed + 3 * slot_bytes <= kSmemOneCta) return 3;
      if (fixed + 2 * slot_bytes <= kSmemOneCta) return 2;
      return 0;
    };
    if (vec == 16 && g->batch > 0) {
      // TMA: one [nch][rows][pitch] box per image the tile touches; everything outside the tensor reads as zero
      const int ro_seg = ro < prm.Ho ? ro : prm.Ho;
      const int rows = (ro_seg - 1) * g->stride + ks;
      const int64_t seg = ((static_cast<int64_t>(nch) * rows * pitch + 127) / 128) * 128;
      const int slots_t = pick_slots(ni * seg);
      CUtensorMap tm;
      const uint64_t Wp = static_cast<uint64_t>(prm.Wp);
      const uint64_t dims[4] = {static_cast<uint64_t>(g->w), static_cast<uint64_t>(g->h), static_cast<uint64_t>(g->c_in),
                                static_cast<uint64_t>(g->batch)};
      const uint64_t strides[3] = {Wp, Wp * g->h, Wp * g->h * g->c_in};
      const uint32_t box[4] = {static_cast<uint32_t>(pitch), static_cast<uint32_t>(rows), static_cast<uint32_t>(nch), 1u};
      if (slots_t && pitch <= 256 && rows <= 256 && encode_tmap(&tm, input_nchw, 4, dims, strides, box)) {
        prm.use_tma = 1;
        prm.ring_slots = slots_t; prm.slot_bytes = static_cast<int32_t>(ni * seg); prm.seg_bytes = static_cast<int32_t>(seg);
        prm.box_bytes = nch * rows * pitch;
        prm.halo_rows = rows; prm.halo_pitch = pitch; prm.halo_lpad = lpad; prm.halo_nch = nch; prm.halo_vec = 16;
        return launch_tc(&plan->p, prm, ks == 3 ? accel::kModeConv3 : accel::kModeConv7,
                         fixed + slots_t * static_cast<int>(ni * seg), st, &tm);
      }
    }
    const int64_t slot = ((static_cast<int64_t>(nch) * hr * pitch + 15) / 16) * 16;
    const int slots = pick_slots(slot);
    if (slots && ro < accel::kMaxOutRows && hr <= accel::kMaxHaloRows) {
      prm.ring_slots = slots; prm.slot_bytes = static_cast<int32_t>(slot);
      prm.halo_rows = hr; prm.halo_pitch = pitch; prm.halo_lpad = lpad; prm.halo_nch = nch;
      prm.halo_vec = vec; prm.halo_ipr = ipr;
      prm.ipr_magic = ipr > 1 ? static_cast<uint32_t>(((1ull << 32) + ipr - 1) / ipr) : 0u;
      return launch_tc(&plan->p, prm, ks == 3 ? accel::kModeConv3 : accel::kModeConv7,
                       fixed + slots * static_cast<int>(slot), st);
    }
  }
  return launch_tc(&plan->p, prm, accel::kModeDirect, accel::kSmemRing, st);
}

int accel_conv_bsr_i8_dual(const accel_plan* plan, const accel_plan* plan_ds, const int8_t* input_nchw,
                           const accel_conv_geom* g, const accel_epilogue* epi, void* out, const accel_epilogue* epi_ds,
                           void* out_ds, const accel_out_layout* layout, accel_stream_t stream) {
  if (!plan || !plan_ds || !g || !epi || !epi_ds) return fail(ACCEL_INVALID_CONFIG, "null plan / geometry / epilogue");
  if (!plan->p.uploaded || !plan_ds->p.uploaded) return fail(ACCEL_NOT_READY, "Weights not loaded");
  if (g->ksize != 3 || g->stride != 2 || g->pad != 1)
    return fail(ACCEL_INVALID_CONFIG, "the dual call is a 3x3 / stride 2 / pad 1 convolution plus the 1x1 / stride 2 one");
  int rc = check_epilogue(epi, layout, out, plan->p.nbr * accel::kBlock);
  if (rc) return rc;
  rc = check_epilogue(epi_ds, layout, out_ds, plan_ds->p.nbr * accel::kBlock);
  if (rc) return rc;
  if (g->batch > 0 && input_nchw && g->c_in > 0 && g->h > 0 && g->w > 0) {
    rc = try_conv_ws(plan, input_nchw, g, epi, out, layout, static_cast<cudaStream_t>(stream), plan_ds, epi_ds, out_ds);
    if (rc != kWsNotApplicable) return rc;
  }
  // not fusable for these tensors: the two convolutions one after the other
  rc = accel_conv_bsr_i8(plan, input_nchw, g, epi, out, layout, stream);
  if (rc) return rc;
  accel_conv_geom g1 = *g;
  g1.ksize = 1; g1.pad = 0;
  return accel_conv_bsr_i8(plan_ds, input_nchw, &g1, epi_ds, out_ds, layout, stream);
}

int accel_conv_pool_bsr_i8(const accel_plan* plan, const int8_t* input_nchw, const accel_conv_geom* g, const accel_epilogue* epi,
                           int32_t pool, int32_t pool_stride, int32_t pool_pad, int8_t* out, int32_t out_pitch, accel_stream_t stream) {
  if (!plan || !g || !epi || !out) return fail(ACCEL_INVALID_CONFIG, "null plan / geometry / epilogue / output");
  if (!plan->p.uploaded) return fail(ACCEL_NOT_READY, "Weights not loaded");
  const WsState& W = plan->ws;
  // the fused kernel exists for the ResNet stem shape family only; anything else is ACCEL_ILLEGAL_COMMAND and the caller
  // runs accel_conv_bsr_i8 + accel_maxpool_i8 (no scratch tensor is owned by this library)
  const bool ok = !g_no_ws && W.ready && W.taps == 49 && g->ksize == 7 && g->stride == 2 && g->pad == 3 && pool == 3 &&
                  pool_stride == 2 && pool_pad == 1 && g->c_in == W.c_in && epi->n_channels == W.c_out &&
                  (epi->flags & ACCEL_OUT_I8) && !epi->chan_absmax && !epi->residual && !(epi->flags & ACCEL_RELU_OUT) &&
                  epi->chan_scale && g->batch > 0 && g->w % 32 == 0 && g->w <= 224 && g->h % 4 == 0 &&
                  g->c_in * 7 * (g->w / 32) <= 160;
  if (!ok) return fail(ACCEL_ILLEGAL_COMMAND, "no fused convolution + max-pool kernel for this geometry");
  const int64_t in_pitch = g->in_row_pitch > 0 ? g->in_row_pitch : g->w;
  const int Hc = g->h / 2, Wc = g->w / 2, Hp = Hc / 2, Wp = Wc / 2;
  if (out_pitch == 0) out_pitch = Wp;
  if ((in_pitch & 15) || (reinterpret_cast<uintptr_t>(input_nchw) & 15) || (out_pitch & 3) || out_pitch < ((Wp + 3) & ~3) ||
      (reinterpret_cast<uintptr_t>(out) & 3) || static_cast<int64_t>(g->batch) * W.c_out * Hp * out_pitch >= (1ll << 40))
    return fail(ACCEL_ILLEGAL_COMMAND, "fused convolution + max-pool needs 16-byte aligned input rows and 4-byte aligned output rows");
  std::call_once(g_attr_once, set_kernel_attrs);
  if (g_attr_err != cudaSuccess) return cuda_fail(g_attr_err, "cudaFuncSetAttribute(smem)");
  accel::StemParams p;
  std::memset(&p, 0, sizeof(p));
  p.C = g->c_in; p.H = g->h; p.W = g->w; p.B = g->batch;
  p.Hc = Hc; p.Wc = Wc; p.Hp = Hp; p.Wp = Wp;
  p.c_out = W.c_out; p.in_pitch = static_cast<int32_t>(in_pitch); p.out_pitch = out_pitch;
  p.n_pairs = (g->batch + 1) / 2;
  // an item is a strip of G pooled rows of an image pair = 2G + 1 conv rows (2G for the first strip); pick the G with the
  // fewest conv rows on the busiest SM
  int best_g = 1;
  int64_t best_cost = INT64_MAX;
  for (int G = 1; G <= Hp; ++G) {
    const int64_t strips = (Hp + G - 1) / G, items = static_cast<int64_t>(p.n_pairs) * strips;
    const int64_t cost = ((items + sm_count() - 1) / sm_count()) * (2 * G + 1);
    if (cost < best_cost) { best_cost = cost; best_g = G; }
  }
  if (g_stem_rows > 0) best_g = g_stem_rows < Hp ? g_stem_rows : Hp;
  p.G = best_g;
  p.n_strips = (Hp + best_g - 1) / best_g;
  const int64_t n_items = static_cast<int64_t>(p.n_pairs) * p.n_strips;
  if (n_items > INT_MAX) return fail(ACCEL_INVALID_CONFIG, "too many rows");
  p.n_items = static_cast<int32_t>(n_items);
  p.d_strips = accel::make_fastdiv(static_cast<uint32_t>(p.n_strips));
  p.wblob = W.blob; p.epi = *epi; p.out = out;
  p.chan_stride = Hp * out_pitch;
  p.image_stride = static_cast<int64_t>(W.c_out) * p.chan_stride;
  p.x = input_nchw;
  p.timeline = g_timeline;
  p.raw_slot_bytes = (g->c_in * 7 * (16 + g->w + 32) + 127) & ~127;
  if (g->c_in * 7 * (g->w / 32 + 1) > 6 * 32 || accel::kStSmemFixed + accel::kStRawSlots * p.raw_slot_bytes > kSmemWs)
    return fail(ACCEL_ILLEGAL_COMMAND, "fused convolution + max-pool: input rows too long for the raw ring");
  p.dbg = g_dbg_flags;
  const int ctas = n_items < sm_count() ? static_cast<int>(n_items) : sm_count();
  const int smem = accel::kStSmemFixed + accel::kStRawSlots * p.raw_slot_bytes;
  cudaError_t e = launch_overlapped(accel::stem_ws_kernel, static_cast<unsigned>(ctas), accel::kStThreads, smem,
                                    static_cast<cudaStream_t>(stream), p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "stem_ws_kernel launch");
  ++g_ws_launches;
  return ACCEL_OK;
}

int accel_bsr_gemm_generic(const int8_t* act, int64_t M, int64_t K, int64_t lda, const int32_t* row_ptr,
                           const int32_t* col_idx, const int8_t* blocks, int32_t n_block_rows, int32_t block_h,
                           int32_t block_w, int32_t orient, int64_t n_out, int32_t* out, int64_t ldo,
                           accel_stream_t stream) {
  if (M < 0 || K < 0 || n_out < 0 || block_h <= 0 || block_w <= 0 || n_block_rows < 0 || ldo < n_out)
    return fail(ACCEL_INVALID_CONFIG, "bad shape");
  if (M == 0 || n_out == 0) return ACCEL_OK;
  if (!act || !row_ptr || !out) return fail(ACCEL_INVALID_CONFIG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(M * n_out, 256, 16);
  if (orient == 0)
    accel::bsr_gemm_generic_b_kernel<<<grid, 256, 0, st>>>(act, M, K, lda, row_ptr, col_idx, blocks, n_block_rows,
                                                           block_h, block_w, n_out, out, ldo);
  else
    accel::bsr_gemm_generic_a_kernel<<<grid, 256, 0, st>>>(act, M, K, lda, row_ptr, col_idx, blocks, n_block_rows,
                                                           block_h, block_w, n_out, out, ldo);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

// ----------------------------------------------------------------------------------- packer
static int grid_dims(int64_t rows, int64_t cols, int32_t bh, int32_t bw, int32_t* nbr, int32_t* nbc) {
  if (rows < 0 || cols < 0 || bh <= 0 || bw <= 0) return fail(ACCEL_INVALID_CONFIG, "bad matrix / block shape");
  *nbr = static_cast<int32_t>((rows + bh - 1) / bh);
  *nbc = static_cast<int32_t>((cols + bw - 1) / bw);
  return ACCEL_OK;
}

int accel_block_l1_i8(const int8_t* w, int64_t rows, int64_t cols, int64_t ld, int32_t block, int32_t* l1_out,
                      accel_stream_t stream) {
  int32_t nbr, nbc;
  if (int rc = grid_dims(rows, cols, block, block, &nbr, &nbc)) return rc;
  if (!nbr || !nbc) return ACCEL_OK;
  accel::block_l1_i8_kernel<<<grid_for(static_cast<int64_t>(nbr) * nbc * 32, 256), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(w, rows, cols, ld, block, nbr, nbc, l1_out);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_block_l2_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, int32_t bh, int32_t bw, float* l2_out,
                       accel_stream_t stream) {
  int32_t nbr, nbc;
  if (int rc = grid_dims(rows, cols, bh, bw, &nbr, &nbc)) return rc;
  if (!nbr || !nbc) return ACCEL_OK;
  accel::block_l2_f32_kernel<<<grid_for(static_cast<int64_t>(nbr) * nbc * 32, 256), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(w, rows, cols, ld, bh, bw, nbr, nbc, l2_out);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_bsr_scan(const uint8_t* keep, int32_t nbr, int32_t nbc, int32_t* row_ptr, int32_t* slot,
                   accel_stream_t stream) {
  if (nbr < 0 || nbc < 0 || !row_ptr) return fail(ACCEL_INVALID_CONFIG, "bad block grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nbr == 0 || nbc == 0) {
    CU(cudaMemsetAsync(row_ptr, 0, sizeof(int32_t) * (static_cast<size_t>(nbr) + 1), st));
    return ACCEL_OK;
  }
  if (!keep || !slot) return fail(ACCEL_INVALID_CONFIG, "null keep / slot");
  // slot[0..nbr) (it has nbr*nbc >= nbr entries) holds the per-row counts until the scan consumed them
  accel::bsr_row_count_kernel<<<grid_for(static_cast<int64_t>(nbr) * 32, 256), 256, 0, st>>>(keep, nbr, nbc, slot);
  CU(cudaGetLastError());
  accel::bsr_row_scan_kernel<<<1, 1024, 0, st>>>(slot, nbr, row_ptr);
  CU(cudaGetLastError());
  if (nbc > 0) {
    accel::bsr_slot_kernel<<<grid_for(static_cast<int64_t>(nbr) * 32, 256), 256, 0, st>>>(keep, nbr, nbc, row_ptr, slot);
    CU(cudaGetLastError());
  }
  return ACCEL_OK;
}

int accel_bsr_gather_i8(const int8_t* w, int64_t rows, int64_t cols, int64_t ld, int32_t block, const int32_t* slot,
                        int32_t nbr, int32_t nbc, int32_t* col_idx, int8_t* blocks, accel_stream_t stream) {
  if (!nbr || !nbc) return ACCEL_OK;
  accel::bsr_gather_i8_kernel<<<grid_for(static_cast<int64_t>(nbr) * nbc * 32, 256), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(w, rows, cols, ld, block, slot, nbr, nbc, col_idx,
                                                                     blocks);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_bsr_gather_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, int32_t bh, int32_t bw,
                         const int32_t* slot, int32_t nbr, int32_t nbc, int32_t* col_idx, float* blocks,
                         accel_stream_t stream) {
  if (!nbr || !nbc) return ACCEL_OK;
  accel::bsr_gather_f32_kernel<<<grid_for(static_cast<int64_t>(nbr) * nbc * 32, 256), 256, 0,
                                 static_cast<cudaStream_t>(stream)>>>(w, rows, cols, ld, bh, bw, slot, nbr, nbc,
                                                                      col_idx, blocks);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_quantize_rows_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, const float* scales, int8_t* q,
                            accel_stream_t stream) {
  if (rows <= 0 || cols <= 0) return ACCEL_OK;
  accel::quantize_rows_f32_kernel<<<grid_for(rows * cols, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, rows, cols, ld, scales, q);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_symmetric_scales_f32(const float* absmax, int64_t n, float* scales, accel_stream_t stream) {
  if (n <= 0) return ACCEL_OK;
  accel::symmetric_scales_f32_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(absmax, n, scales);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_row_absmax_f32(const float* w, int64_t rows, int64_t cols, int64_t ld, float* absmax,
                         accel_stream_t stream) {
  if (rows <= 0) return ACCEL_OK;
  accel::row_absmax_f32_kernel<<<grid_for(rows * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, rows, cols,
                                                                                                       ld, absmax);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

// ----------------------------------------------------------------------------------- epilogue pieces / pools
int accel_requant_i32_i8(const int32_t* acc, int8_t* out, int64_t n_outer, int64_t n_chan, int64_t n_inner,
                         const float* chan_scale, const int32_t* bias, int32_t relu, unsigned long long* sat_count,
                         accel_stream_t stream) {
  const int64_t total = n_outer * n_chan * n_inner;
  if (total <= 0) return ACCEL_OK;
  if (!chan_scale) return fail(ACCEL_INVALID_CONFIG, "chan_scale required");
  accel::requant_i32_i8_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, out, n_outer, n_chan, n_inner, chan_scale, bias, relu, sat_count);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_subsample2_i8(const int8_t* in, int64_t planes, int32_t h, int32_t w, int32_t in_pitch, int8_t* out, int32_t out_pitch,
                        accel_stream_t stream) {
  if (planes <= 0 || h <= 0 || w <= 0) return ACCEL_OK;
  if (!in || !out) return fail(ACCEL_INVALID_CONFIG, "null buffer");
  const int Ho = (h + 1) / 2, Wo = (w + 1) / 2;
  if (in_pitch < w || out_pitch < Wo) return fail(ACCEL_INVALID_CONFIG, "row pitch shorter than the row");
  const int quads = (Wo + 3) / 4;
  const bool vec = !(in_pitch & 7) && !(out_pitch & 3) && 8 * quads <= in_pitch && !(reinterpret_cast<uintptr_t>(in) & 7) &&
                   !(reinterpret_cast<uintptr_t>(out) & 3);
  if (vec)
    accel::subsample2_i8_vec_kernel<<<grid_for(planes * Ho * quads, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, planes, in_pitch, h, Ho, Wo, out_pitch);
  else
    accel::subsample2_i8_kernel<<<grid_for(planes * Ho * Wo, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, planes, in_pitch, h, Ho, Wo, out_pitch);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_relu_i8(int8_t* data, int64_t n, accel_stream_t stream) {
  if (n <= 0) return ACCEL_OK;
  if (!data) return fail(ACCEL_INVALID_CONFIG, "null buffer");
  accel::relu_i8_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(data, n, 127);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_relu6_i8(int8_t* data, int64_t n, float scale, accel_stream_t stream) {
  if (n <= 0) return ACCEL_OK;
  if (!data) return fail(ACCEL_INVALID_CONFIG, "null buffer");
  // the reference's threshold, evaluated the way it is written there (golden_models.cpp:326): a float -> int8 cast of
  // 6.0f / scale (truncation toward zero; like the reference this is only meaningful while 6 / scale fits an int8)
  const std::int8_t max_val = static_cast<std::int8_t>(6.0f / scale);
  accel::relu_i8_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(data, n, static_cast<int32_t>(max_val));
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_relu_i32(int32_t* data, int64_t n, accel_stream_t stream) {
  if (n <= 0) return ACCEL_OK;
  if (!data) return fail(ACCEL_INVALID_CONFIG, "null buffer");
  accel::relu_i32_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(data, n);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_gemm_bsr_int8_fp32(const int8_t* A, int64_t M, int64_t lda, const int32_t* indptr, const int32_t* indices,
                             const void* data, const void* scales_div, int32_t div_f64, const int32_t* blk_row, int64_t nnz,
                             int32_t n_block_rows, int32_t block, int64_t K, int64_t N, double scale_a, int32_t scale_a_f64,
                             const void* scales, int32_t scales_f64, int32_t n_scales, int8_t* q_scratch, float* C, int64_t ldc,
                             accel_stream_t stream) {
  if (M < 0 || K < 0 || N < 0 || block <= 0 || n_block_rows < 0 || ldc < N || n_scales <= 0)
    return fail(ACCEL_INVALID_CONFIG, "bad shape");
  if (K % block) return fail(ACCEL_INVALID_CONFIG, "K must be a multiple of the block size (the reference's A-slice @ block^T needs whole blocks)");
  if (M == 0 || N == 0) return ACCEL_OK;
  if (!A || !indptr || !C || !scales || (nnz > 0 && (!data || !scales_div || !indices || !blk_row || !q_scratch)))
    return fail(ACCEL_INVALID_CONFIG, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nnz > 0) {
    const int64_t tot = nnz * block * block;
    // block / scale is float64 as soon as either operand is (NumPy promotion): the wrapper passes both in that dtype
    if (div_f64)
      accel::fp32compat_requant_kernel<double><<<grid_for(tot, 256), 256, 0, st>>>(static_cast<const double*>(data), blk_row, nnz, block,
                                                                                 block, static_cast<const double*>(scales_div), n_scales, q_scratch);
    else
      accel::fp32compat_requant_kernel<float><<<grid_for(tot, 256), 256, 0, st>>>(static_cast<const float*>(data), blk_row, nnz, block,
                                                                                block, static_cast<const float*>(scales_div), n_scales, q_scratch);
    CU(cudaGetLastError());
  }
  accel::fp32compat_gemm_kernel<<<grid_for(M * N, 256), 256, 0, st>>>(A, M, lda, indptr, indices, q_scratch, n_block_rows, block, K, N,
                                                                       scale_a, scale_a_f64, scales, scales_f64, n_scales, C, ldc);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_add_residual_i8(const int8_t* main_, const int8_t* res, int8_t* out, int64_t n, float s_main, float s_res,
                          float s_out, accel_stream_t stream) {
  if (n <= 0) return ACCEL_OK;
  accel::add_residual_i8_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(main_, res, out, n,
                                                                                                 s_main, s_res, s_out);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_maxpool_i8(const int8_t* x, int8_t* out, int64_t n_planes, int32_t h, int32_t w, int32_t pool,
                     int32_t stride, int32_t pad, int32_t in_pitch, int32_t out_pitch, accel_stream_t stream) {
  if (pool <= 0 || stride <= 0 || pad < 0 || h + 2 * pad < pool || w + 2 * pad < pool)
    return fail(ACCEL_INVALID_CONFIG, "bad pooling geometry");
  const int32_t Ho = (h + 2 * pad - pool) / stride + 1, Wo = (w + 2 * pad - pool) / stride + 1;
  if (in_pitch == 0) in_pitch = w;
  if (out_pitch == 0) out_pitch = Wo;
  if (in_pitch < w || out_pitch < Wo) return fail(ACCEL_INVALID_CONFIG, "row pitch smaller than the row");
  const int64_t total = n_planes * Ho * ((Wo + 3) / 4);
  if (total <= 0) return ACCEL_OK;
  if (pool == 3 && stride == 2 && pad == 1 && (in_pitch & 15) == 0 && (out_pitch & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int wq = (Wo + 15) / 16;
    const int64_t total16 = n_planes * Ho * wq;
    if (total16 < (1ll << 31)) {
      accel::maxpool3x3s2_i8_x16_kernel<<<grid_for(total16, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          x, out, static_cast<uint32_t>(total16), h, w, Ho, Wo, in_pitch, out_pitch,
          accel::make_fastdiv(static_cast<uint32_t>(wq)), accel::make_fastdiv(static_cast<uint32_t>(Ho)));
      CU(cudaGetLastError());
      return ACCEL_OK;
    }
  }
  if (pool == 3 && stride == 2 && pad == 1 && (in_pitch & 7) == 0 && (out_pitch & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    accel::maxpool3x3s2_i8_kernel<<<grid_for(total, 256, 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, out, n_planes, h, w, Ho, Wo, in_pitch, out_pitch);
    CU(cudaGetLastError());
    return ACCEL_OK;
  }
  accel::maxpool_i8_kernel<<<grid_for(total, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, out, n_planes, h, w, pool, stride, pad, Ho, Wo, in_pitch, out_pitch);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

int accel_avgpool_i8(const int8_t* x, int8_t* out, int64_t n_planes, int32_t h, int32_t w, int32_t in_pitch,
                     accel_stream_t stream) {
  if (n_planes <= 0) return ACCEL_OK;
  if (in_pitch == 0) in_pitch = w;
  if (in_pitch < w) return fail(ACCEL_INVALID_CONFIG, "row pitch smaller than the row");
  if (w <= 16 && (in_pitch & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n_planes < (1ll << 28)) {
    accel::avgpool_i8_rows16_kernel<<<static_cast<unsigned>((n_planes * 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, out, static_cast<uint32_t>(n_planes), h, w, in_pitch);
    CU(cudaGetLastError());
    return ACCEL_OK;
  }
  accel::avgpool_i8_kernel<<<grid_for(n_planes * 8, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, n_planes,
                                                                                                        h, w, in_pitch);
  CU(cudaGetLastError());
  return ACCEL_OK;
}

}  // extern "C"
