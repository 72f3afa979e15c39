"""Torch-facing wrappers over the C ABI: device memory and streams come from PyTorch, every
byte of arithmetic happens in libaccel_b200.so.  No fallbacks: without CUDA these raise."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import AcceleratorError, ConvGeom, Epilogue, OutLayout, check

BLOCK = 14
# GEMMs with at least this many activation rows build the dense-equivalent layout of csrc/gemm_ws.cuh on first use
GEMM_WS_MIN_ROWS = int(__import__("os").environ.get("ACCEL_GEMM_WS_MIN_ROWS", "128"))


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise AcceleratorError(_lib.INIT_FAILED, "no CUDA device: resnet_accel_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def to_device(a, dtype: torch.dtype, device=None) -> torch.Tensor:
    """numpy / torch (any device) -> contiguous CUDA tensor of `dtype` (zero-copy when already there)."""
    device = device or _require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint16:
            a = a.astype(np.int32)
        elif a.dtype == np.uint32:
            a = a.astype(np.int64)
        t = torch.from_numpy(a)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=False).contiguous()


def alloc_padded(shape, device=None, align: int = 16) -> torch.Tensor:
    """int8 NCHW activation tensor whose rows start ``align``-byte aligned: a ``[..., :W]`` view of a zeroed
    ``[B, C, H, Wp]`` buffer (Wp = W rounded up).  The conv loader streams such rows with 16-byte cp.async."""
    device = device or _require_cuda()
    *lead, W = shape
    Wp = -(-W // align) * align
    return torch.zeros((*lead, Wp), dtype=torch.int8, device=device)[..., :W]


def _row_pitch(t: torch.Tensor) -> Optional[int]:
    """Row pitch (elements) of an NCHW tensor that is dense except for padded rows, else None."""
    B, Cn, H, W = t.shape
    if t.stride(3) != 1 and W > 1:
        return None
    Wp = t.stride(2) if H > 1 else (t.stride(1) // max(H, 1) if Cn > 1 else W)
    if Wp < W:
        return None
    if (Cn > 1 and t.stride(1) != H * Wp) or (B > 1 and t.stride(0) != Cn * H * Wp):
        return None
    return int(Wp)


class BsrPlan:
    """A BSR weight matrix (Convention B, 14x14) resident on the GPU in MMA-tile form.

    Mirrors what ``AccelDriver.load_sparse_weights`` does for the FPGA (sw/host/accel.py:177-236):
    takes ``row_ptr`` / ``col_idx`` / ``blocks[nnz,14,14]`` exactly as the exporters write them.
    """

    def __init__(self, row_ptr, col_idx, blocks, n_block_cols: Optional[int] = None, group_rows: int = 0):
        self.device = _require_cuda()
        L = _lib.lib()
        rp = np.ascontiguousarray(_host(row_ptr), dtype=np.int32)
        ci = np.ascontiguousarray(_host(col_idx), dtype=np.int32)
        if rp.ndim != 1 or rp.size < 1:
            raise AcceleratorError(_lib.INVALID_CONFIG, "row_ptr must be 1-D with >= 1 entry")
        self.n_block_rows = int(rp.size - 1)
        if n_block_cols is None:
            n_block_cols = int(ci.max()) + 1 if ci.size else 0
        self.n_block_cols = int(n_block_cols)
        if int(rp[-1]) != ci.size:
            raise AcceleratorError(_lib.INVALID_CONFIG, f"col_idx size mismatch: expected {int(rp[-1])}, got {ci.size}")
        blk = to_device(blocks, torch.int8, self.device) if not _is_cuda(blocks) else blocks.contiguous()
        if blk.numel() != ci.size * BLOCK * BLOCK:
            raise AcceleratorError(_lib.INVALID_CONFIG,
                                   f"data size mismatch: expected {ci.size * 196}, got {blk.numel()}")
        handle = C.c_void_p()
        ws_bytes = C.c_size_t()
        check(L.accel_plan_create(rp.ctypes.data, ci.ctypes.data if ci.size else None, self.n_block_rows,
                                  self.n_block_cols, BLOCK, int(group_rows), C.byref(handle), C.byref(ws_bytes)))
        self._h = handle
        self.workspace = torch.empty(max(int(ws_bytes.value), 256) + 256, dtype=torch.uint8, device=self.device)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_ptr = self.workspace.data_ptr() + off
        check(L.accel_plan_upload(self._h, _ptr(blk) if ci.size else None, self._ws_ptr, int(ws_bytes.value), _stream()))
        self.num_blocks = int(L.accel_plan_num_blocks(self._h))
        self.num_mma = int(L.accel_plan_num_mma(self._h))
        self.row_ptr, self.col_idx = rp, ci
        self._blocks = blk                      # kept for later re-layouts (conv_ws)
        self._ws_conv = {}                      # (c_in, c_out, ksize) -> workspace tensor (None: no such path)
        self._ws_gemm = None                    # dense-equivalent GEMM layout: None = not built yet, False = no such path
        self.use_gemm_ws = True                 # False: GEMMs of this plan stay on the gather kernel (csrc/bsr_tcp.cuh)
        self._row_l1_max = None                 # largest sum of |w| over one output channel's stored weights (lazy)
        self._bound_cache = {}                  # bias identity -> acc_bound

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().accel_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def _prepare_conv_ws(self, c_in: int, c_out: int, ksize: int) -> None:
        """Weight-stationary layout for 3x3 stride-1 convolutions (csrc/conv_ws.cuh), built once per geometry."""
        key = (int(c_in), int(c_out), int(ksize))
        if key in self._ws_conv:
            return
        # the plan holds one prepared geometry at a time.  The C side is told to forget the old layout BEFORE its buffer
        # can go back to the allocator (accel_plan_conv_ws_release), and the old buffer is only dropped afterwards.
        L = _lib.lib()
        nbytes = C.c_size_t()
        check(L.accel_plan_conv_ws_bytes(self._h, key[0], key[1], key[2], C.byref(nbytes)))
        old = self._ws_conv
        L.accel_plan_conv_ws_release(self._h)
        if nbytes.value == 0:
            self._ws_conv = {key: None}
            del old
            return
        buf = torch.empty(int(nbytes.value) + 1024, dtype=torch.uint8, device=self.device)
        ptr = buf.data_ptr() + (-buf.data_ptr()) % 1024
        check(L.accel_plan_conv_ws_prepare(self._h, _ptr(self._blocks) if self.col_idx.size else None, key[0], key[1], key[2],
                                           ptr, int(nbytes.value), _stream()))
        self._ws_conv = {key: buf}
        del old

    def _prepare_gemm_ws(self) -> None:
        """Dense-equivalent layout for GEMMs (csrc/gemm_ws.cuh), built once per plan on the first large GEMM."""
        if self._ws_gemm is not None:
            return
        L = _lib.lib()
        nbytes = C.c_size_t()
        check(L.accel_plan_gemm_ws_bytes(self._h, C.byref(nbytes)))
        if nbytes.value == 0:
            self._ws_gemm = False
            return
        buf = torch.empty(int(nbytes.value) + 256, dtype=torch.uint8, device=self.device)
        ptr = buf.data_ptr() + (-buf.data_ptr()) % 256
        check(L.accel_plan_gemm_ws_prepare(self._h, _ptr(self._blocks) if self.col_idx.size else None, ptr, int(nbytes.value),
                                           _stream()))
        self._ws_gemm = buf

    @property
    def n_out_padded(self) -> int:
        return self.n_block_rows * BLOCK

    def acc_bound(self, bias=None, may_sync: bool = True) -> int:
        """``accel_epilogue.acc_bound``: no int8 input can drive ``|accumulator + bias|`` past
        ``128 * max_channel(sum |w|) + max |bias|``.  Below 2**22 the weight-stationary kernels run their conversion-free
        epilogue (csrc/conv_ws.cuh, ws_epi16).  One device reduction per plan and per bias tensor, cached; with ``may_sync``
        False (stream capture) an uncached bound is reported as 0 = unknown."""
        if self._row_l1_max is None:
            if not may_sync:
                return 0
            if self.col_idx.size == 0:
                self._row_l1_max = 0
            else:
                l1 = self._blocks.view(-1, BLOCK, BLOCK).to(torch.int32).abs().sum(2)            # [nnz, 14]
                owner = torch.repeat_interleave(torch.arange(self.n_block_rows, device=self.device),
                                                torch.as_tensor(np.diff(self.row_ptr), device=self.device))
                rows = torch.zeros((self.n_block_rows, BLOCK), dtype=torch.int64, device=self.device).index_add_(0, owner, l1)
                self._row_l1_max = int(rows.max().item())
        if bias is None:
            bmax = 0
        elif _is_cuda(bias):
            key = (bias.data_ptr(), bias.numel(), bias._version)
            if key not in self._bound_cache:
                if not may_sync:
                    return 0
                self._bound_cache[key] = int(bias.to(torch.int64).abs().max().item()) if bias.numel() else 0
            bmax = self._bound_cache[key]
        else:
            b = np.asarray(_host(bias), dtype=np.int64)
            bmax = int(np.abs(b).max()) if b.size else 0
        return min(128 * self._row_l1_max + bmax, 2**31 - 1)

    # ------------------------------------------------------------------------------------
    def _epilogue(self, out_kind: str, n_channels: int, chan_scale, bias, relu: bool, residual, res_scales,
                  sat_count, chan_absmax, relu_out: bool = False) -> Tuple[Epilogue, list]:
        keep = []
        flags = {"i8": _lib.OUT_I8, "i32": _lib.OUT_I32, "f32": _lib.OUT_F32}[out_kind] | (_lib.RELU if relu else 0)
        if relu_out:
            flags |= _lib.RELU_OUT
        e = Epilogue()
        e.flags, e.n_channels = flags, int(n_channels)
        for name, val, dt in (("chan_scale", chan_scale, torch.float32), ("bias", bias, torch.int32),
                              ("residual", residual, torch.int8)):
            if val is not None:
                if name == "residual" and _is_cuda(val) and val.dtype == dt:
                    t = val                                  # may have padded rows: same strides as the output
                else:
                    t = val if _is_cuda(val) and val.dtype == dt and val.is_contiguous() else to_device(val, dt, self.device)
                if name != "residual" and t.numel() < n_channels:
                    raise AcceleratorError(_lib.INVALID_CONFIG, f"{name} has {t.numel()} entries, need {n_channels}")
                keep.append(t)
                setattr(e, name, t.data_ptr())
        s = res_scales or (1.0, 1.0, 1.0)
        e.res_scale_main, e.res_scale_res, e.res_scale_out = float(s[0]), float(s[1]), float(s[2])
        if sat_count is not None:
            e.sat_count = sat_count.data_ptr()
        if chan_absmax is not None:
            e.chan_absmax = chan_absmax.data_ptr()
        if out_kind == "i8":                                 # cached after the first call with this bias tensor
            e.acc_bound = self.acc_bound(bias, may_sync=not torch.cuda.is_current_stream_capturing())
        return e, keep

    def gemm(self, x: torch.Tensor, out_kind: str = "i32", n_channels: Optional[int] = None, chan_scale=None,
             bias=None, relu: bool = False, residual=None, res_scales=None, out: Optional[torch.Tensor] = None,
             sat_count: Optional[torch.Tensor] = None, chan_absmax: Optional[torch.Tensor] = None,
             relu_out: bool = False, out_t: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Y = epilogue(X @ W^T).  x int8 [M, K] (CUDA, row stride may exceed K) -> [M, n_channels].

        ``out_t``: write the TRANSPOSED result instead, into a caller-owned ``[n_channels, M]`` tensor whose rows may be
        strided (a slice of a larger buffer): channel-major output is what the block-row-sharded FC hands to the all-gather
        (``parallel.ShardedBsrLinear``).  Returns ``out_t`` in that case."""
        if x.dtype != torch.int8 or x.dim() != 2 or not x.is_cuda:
            raise AcceleratorError(_lib.INVALID_CONFIG, "Activations must be a 2-D INT8 CUDA tensor")
        if x.stride(1) != 1:
            x = x.contiguous()
        M, K = x.shape
        n_channels = self.n_out_padded if n_channels is None else int(n_channels)
        dt = {"i8": torch.int8, "i32": torch.int32, "f32": torch.float32}[out_kind]
        if out_t is not None:
            if out is not None or residual is not None:
                raise AcceleratorError(_lib.INVALID_CONFIG, "out_t excludes out / residual")
            _check_out(out_t, (n_channels, M), dt)
            if M > 1 and out_t.stride(1) != 1:
                raise AcceleratorError(_lib.INVALID_CONFIG, "out_t rows must be contiguous")
            e, keep = self._epilogue(out_kind, n_channels, chan_scale, bias, relu, None, None, sat_count, chan_absmax, relu_out)
            if (self.use_gemm_ws and M >= GEMM_WS_MIN_ROWS and x.stride(0) % 16 == 0 and x.data_ptr() % 16 == 0
                    and chan_absmax is None):
                self._prepare_gemm_ws()
            lay = OutLayout(max(M, 1), 0, out_t.stride(0) if n_channels > 1 else max(M, 1), 1, 0, 0)
            check(_lib.lib().accel_bsr_gemm_i8(self._h, _ptr(x), M, K, x.stride(0) if M else K, C.byref(e), _ptr(out_t),
                                               C.byref(lay), _stream()))
            return out_t
        if out is None:
            out = torch.empty((M, n_channels), dtype=dt, device=x.device)
        _check_out(out, (M, n_channels), dt)
        if out.dim() != 2 or (out.stride(1) != 1 and n_channels > 1):
            raise AcceleratorError(_lib.INVALID_CONFIG, "output must be 2-D with unit column stride")
        if residual is not None:
            if not (_is_cuda(residual) and residual.dtype == torch.int8):
                residual = to_device(residual, torch.int8, x.device)
            if tuple(residual.shape) != tuple(out.shape):
                raise AcceleratorError(_lib.INVALID_CONFIG, "residual must have the output's shape")
            if residual.stride() != out.stride():
                r2 = torch.empty_strided(tuple(out.shape), tuple(out.stride()), dtype=torch.int8, device=out.device)
                r2.copy_(residual)
                residual = r2
        e, keep = self._epilogue(out_kind, n_channels, chan_scale, bias, relu, residual, res_scales, sat_count,
                                 chan_absmax, relu_out)
        if (self.use_gemm_ws and M >= GEMM_WS_MIN_ROWS and x.stride(0) % 16 == 0 and x.data_ptr() % 16 == 0
                and chan_absmax is None):
            self._prepare_gemm_ws()
        lay = OutLayout(max(M, 1), 0, 1, out.stride(0) if M else n_channels, 0, 0)
        check(_lib.lib().accel_bsr_gemm_i8(self._h, _ptr(x), M, K, x.stride(0) if M else K, C.byref(e), _ptr(out),
                                           C.byref(lay), _stream()))
        return out

    def conv(self, x: torch.Tensor, ksize: int, stride: int, pad: int, c_out: int, out_kind: str = "i8",
             chan_scale=None, bias=None, relu: bool = False, residual=None, res_scales=None,
             out: Optional[torch.Tensor] = None, sat_count: Optional[torch.Tensor] = None,
             chan_absmax: Optional[torch.Tensor] = None, relu_out: bool = False) -> torch.Tensor:
        """Implicit-im2col BSR convolution.  x int8 NCHW (CUDA) -> [B, c_out, Ho, Wo].  Input and output may have
        padded rows (``alloc_padded``); the residual must share the output's strides."""
        if x.dtype != torch.int8 or x.dim() != 4 or not x.is_cuda:
            raise AcceleratorError(_lib.INVALID_CONFIG, "Activations must be a 4-D INT8 CUDA tensor (NCHW)")
        in_pitch = _row_pitch(x)
        if in_pitch is None:
            x = x.contiguous()
            in_pitch = x.shape[3]
        B, Cin, H, W = x.shape
        Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
        dt = {"i8": torch.int8, "i32": torch.int32, "f32": torch.float32}[out_kind]
        if out is None:
            out = torch.empty((B, c_out, Ho, Wo), dtype=dt, device=x.device)
        _check_out(out, (B, c_out, Ho, Wo), dt)
        out_pitch = _row_pitch(out)
        if out_pitch is None:
            raise AcceleratorError(_lib.INVALID_CONFIG, "output must be NCHW, dense or with padded rows")
        if residual is not None:
            if tuple(residual.shape) != tuple(out.shape):
                raise AcceleratorError(_lib.INVALID_CONFIG, "residual shape must equal the output shape")
            # the kernels read the residual with the OUTPUT's layout: bring it to a CUDA int8 tensor first, then give it the
            # output's strides (a dense host array next to a padded-row output would otherwise be read out of bounds)
            if not (_is_cuda(residual) and residual.dtype == torch.int8):
                residual = to_device(residual, torch.int8, x.device)
            if _row_pitch(residual) != out_pitch:
                r2 = torch.empty_strided(tuple(out.shape), tuple(out.stride()), dtype=torch.int8, device=out.device)
                r2.copy_(residual)
                residual = r2
        e, keep = self._epilogue(out_kind, c_out, chan_scale, bias, relu, residual, res_scales, sat_count, chan_absmax,
                                 relu_out)
        if out_kind == "i8" and chan_absmax is None and W <= 62 and (
                (ksize == 3 and stride in (1, 2) and pad == 1) or (ksize == 1 and stride == 1 and pad == 0)):
            self._prepare_conv_ws(Cin, c_out, ksize)
        g = ConvGeom(B, Cin, H, W, ksize, stride, pad, in_pitch)
        P = max(Ho * Wo, 1)
        if out_pitch == Wo:
            lay = OutLayout(P, c_out * Ho * Wo, Ho * Wo, 1, 0, 0)
        else:
            lay = OutLayout(P, c_out * Ho * out_pitch, Ho * out_pitch, 1, Wo, out_pitch)
        check(_lib.lib().accel_conv_bsr_i8(self._h, _ptr(x), C.byref(g), C.byref(e), _ptr(out), C.byref(lay), _stream()))
        return out


def conv_dual(plan: BsrPlan, plan_ds: BsrPlan, x: torch.Tensor, c_out: int, *, chan_scale, chan_scale_ds, bias=None, bias_ds=None,
              relu: bool = True, relu_ds: bool = False, out: Optional[torch.Tensor] = None,
              out_ds: Optional[torch.Tensor] = None, sat_count: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """First convolution of a ResNet stage (3x3 / stride 2 / pad 1, ``plan``) and its downsample branch (1x1 / stride 2,
    ``plan_ds``) from the same input in one call: both int8 NCHW outputs ``[B, c_out, H/2, W/2]`` with the same strides."""
    if x.dtype != torch.int8 or x.dim() != 4 or not x.is_cuda:
        raise AcceleratorError(_lib.INVALID_CONFIG, "Activations must be a 4-D INT8 CUDA tensor (NCHW)")
    in_pitch = _row_pitch(x)
    if in_pitch is None:
        x = x.contiguous()
        in_pitch = x.shape[3]
    B, Cin, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    if out is None:
        out = alloc_padded((B, c_out, Ho, Wo))
    if out_ds is None:
        out_ds = alloc_padded((B, c_out, Ho, Wo))
    out_pitch = _row_pitch(out)
    if out_pitch is None or _row_pitch(out_ds) != out_pitch or tuple(out.shape) != tuple(out_ds.shape):
        raise AcceleratorError(_lib.INVALID_CONFIG, "both outputs must be NCHW with the same shape and row pitch")
    if W <= 62:
        plan._prepare_conv_ws(Cin, c_out, 3)
        plan_ds._prepare_conv_ws(Cin, c_out, 1)
    e, keep = plan._epilogue("i8", c_out, chan_scale, bias, relu, None, None, sat_count, None)
    e2, keep2 = plan_ds._epilogue("i8", c_out, chan_scale_ds, bias_ds, relu_ds, None, None, sat_count, None)
    g = ConvGeom(B, Cin, H, W, 3, 2, 1, in_pitch)
    P = max(Ho * Wo, 1)
    if out_pitch == Wo:
        lay = OutLayout(P, c_out * Ho * Wo, Ho * Wo, 1, 0, 0)
    else:
        lay = OutLayout(P, c_out * Ho * out_pitch, Ho * out_pitch, 1, Wo, out_pitch)
    check(_lib.lib().accel_conv_bsr_i8_dual(plan._h, plan_ds._h, _ptr(x), C.byref(g), C.byref(e), _ptr(out), C.byref(e2),
                                            _ptr(out_ds), C.byref(lay), _stream()))
    return out, out_ds


def conv_pool(plan: BsrPlan, x: torch.Tensor, c_out: int, *, chan_scale, bias=None, relu: bool = True, ksize: int = 7,
              stride: int = 2, pad: int = 3, pool: int = 3, pool_stride: int = 2, pool_pad: int = 1,
              out: Optional[torch.Tensor] = None, sat_count: Optional[torch.Tensor] = None,
              scratch: Optional[torch.Tensor] = None, fused_ok: Optional[bool] = None) -> torch.Tensor:
    """Convolution + ReLU/requant + max-pool (the ResNet stem).  One fused kernel when the library has one for the
    geometry, else the convolution into ``scratch`` (allocated here when missing) followed by the pool.

    The fused kernel pools on the INT32 accumulators, which equals requant-then-pool only when every ``chan_scale[c] > 0``.
    That is a property of ``chan_scale``, not of the plan: callers that know it pass ``fused_ok`` (``BsrLayer.sf_positive``,
    decided on the host when the factors are built); otherwise host-side factors are checked here and device tensors are
    read back once per call (a blocking sync - do not capture such a call into a CUDA graph)."""
    if x.dtype != torch.int8 or x.dim() != 4 or not x.is_cuda:
        raise AcceleratorError(_lib.INVALID_CONFIG, "Activations must be a 4-D INT8 CUDA tensor (NCHW)")
    in_pitch = _row_pitch(x)
    if in_pitch is None:
        x = x.contiguous()
        in_pitch = x.shape[3]
    B, Cin, H, W = x.shape
    Hc, Wc = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    Hp, Wp = (Hc + 2 * pool_pad - pool) // pool_stride + 1, (Wc + 2 * pool_pad - pool) // pool_stride + 1
    if out is None:
        out = alloc_padded((B, c_out, Hp, Wp))
    out_pitch = _row_pitch(out)
    if out_pitch is None:
        raise AcceleratorError(_lib.INVALID_CONFIG, "output must be NCHW, dense or with padded rows")
    fused = ksize == 7 and stride == 2 and pad == 3 and (pool, pool_stride, pool_pad) == (3, 2, 1) and Cin * 7 <= 32 and c_out <= 64
    if fused:
        if fused_ok is None:
            sf_host = chan_scale.detach().cpu().numpy() if isinstance(chan_scale, torch.Tensor) else np.asarray(chan_scale)
            fused_ok = bool((sf_host.reshape(-1)[:c_out] > 0).all())
        fused = bool(fused_ok)
    if fused:
        plan._prepare_conv_ws(Cin, c_out, 7)
        e, keep = plan._epilogue("i8", c_out, chan_scale, bias, relu, None, None, sat_count, None)
        g = ConvGeom(B, Cin, H, W, ksize, stride, pad, in_pitch)
        rc = _lib.lib().accel_conv_pool_bsr_i8(plan._h, _ptr(x), C.byref(g), C.byref(e), pool, pool_stride, pool_pad, _ptr(out),
                                               out_pitch, _stream())
        if rc == 0:
            return out
        if rc != _lib.ILLEGAL_COMMAND:
            check(rc)
    if scratch is None:
        scratch = alloc_padded((B, c_out, Hc, Wc))
    plan.conv(x, ksize, stride, pad, c_out, "i8", chan_scale=chan_scale, bias=bias, relu=relu, out=scratch, sat_count=sat_count)
    return maxpool_i8(scratch, pool, pool_stride, pool_pad, out=out)


def _is_cuda(a) -> bool:
    return isinstance(a, torch.Tensor) and a.is_cuda


def _check_out(out, shape, dtype) -> None:
    """A caller-supplied output buffer must match what the kernel writes (an int8 buffer under out_kind='i32' would overrun)."""
    if not _is_cuda(out) or out.dtype != dtype or tuple(out.shape) != tuple(shape):
        raise AcceleratorError(_lib.INVALID_CONFIG,
                               f"out must be a CUDA {dtype} tensor of shape {tuple(shape)}, got "
                               f"{getattr(out, 'dtype', type(out))} {tuple(getattr(out, 'shape', ()))}")


def _host(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


# ---------------------------------------------------------------------------------------- generic block sizes
def bsr_gemm_generic(x: torch.Tensor, row_ptr, col_idx, blocks, n_out: int, orient: int = 0) -> torch.Tensor:
    """CUDA-core BSR GEMM for any block size; orient 0 = Convention B, 1 = Convention A."""
    dev = _require_cuda()
    x = to_device(x, torch.int8, dev)
    rp = to_device(row_ptr, torch.int32, dev)
    ci = to_device(col_idx, torch.int32, dev)
    blk = to_device(blocks, torch.int8, dev)
    bh, bw = (int(blk.shape[1]), int(blk.shape[2])) if blk.dim() == 3 and blk.shape[0] else (BLOCK, BLOCK)
    M, K = x.shape
    out = torch.empty((M, n_out), dtype=torch.int32, device=dev)
    check(_lib.lib().accel_bsr_gemm_generic(_ptr(x), M, K, x.stride(0) if M else K, _ptr(rp), _ptr(ci) if ci.numel() else None,
                                            _ptr(blk) if blk.numel() else None, rp.numel() - 1, bh, bw, orient, n_out,
                                            _ptr(out), n_out, _stream()))
    return out


# ---------------------------------------------------------------------------------------- packer / pruner (K4)
def block_l1_i8(w: torch.Tensor, block: int = BLOCK) -> torch.Tensor:
    rows, cols = w.shape
    nbr, nbc = -(-rows // block), -(-cols // block)
    out = torch.empty((nbr, nbc), dtype=torch.int32, device=w.device)
    check(_lib.lib().accel_block_l1_i8(_ptr(w), rows, cols, w.stride(0), block, _ptr(out), _stream()))
    return out


def block_l2_f32(w: torch.Tensor, bh: int, bw: int) -> torch.Tensor:
    rows, cols = w.shape
    nbr, nbc = -(-rows // bh), -(-cols // bw)
    out = torch.empty((nbr, nbc), dtype=torch.float32, device=w.device)
    check(_lib.lib().accel_block_l2_f32(_ptr(w), rows, cols, w.stride(0), bh, bw, _ptr(out), _stream()))
    return out


def pack_bsr_i8(w: torch.Tensor, keep: torch.Tensor, block: int = BLOCK):
    """dense int8 [rows, cols] + keep mask [nbr, nbc] -> (row_ptr i32, col_idx i32, blocks i8 [nnz,b,b])."""
    rows, cols = w.shape
    nbr, nbc = keep.shape
    keep8 = keep.to(torch.uint8).contiguous()
    row_ptr = torch.empty(nbr + 1, dtype=torch.int32, device=w.device)
    slot = torch.empty(max(nbr * nbc, 1), dtype=torch.int32, device=w.device)
    check(_lib.lib().accel_bsr_scan(_ptr(keep8), nbr, nbc, _ptr(row_ptr), _ptr(slot), _stream()))
    nnz = int(row_ptr[-1].item())
    col_idx = torch.empty(nnz, dtype=torch.int32, device=w.device)
    blocks = torch.empty((nnz, block, block), dtype=torch.int8, device=w.device)
    if nnz:
        check(_lib.lib().accel_bsr_gather_i8(_ptr(w), rows, cols, w.stride(0), block, _ptr(slot), nbr, nbc,
                                             _ptr(col_idx), _ptr(blocks), _stream()))
    return row_ptr, col_idx, blocks


def pack_bsr_f32(w: torch.Tensor, keep: torch.Tensor, bh: int, bw: int):
    """float32 flavour of :func:`pack_bsr_i8` (generic block shape)."""
    rows, cols = w.shape
    nbr, nbc = keep.shape
    keep8 = keep.to(torch.uint8).contiguous()
    row_ptr = torch.empty(nbr + 1, dtype=torch.int32, device=w.device)
    slot = torch.empty(max(nbr * nbc, 1), dtype=torch.int32, device=w.device)
    check(_lib.lib().accel_bsr_scan(_ptr(keep8), nbr, nbc, _ptr(row_ptr), _ptr(slot), _stream()))
    nnz = int(row_ptr[-1].item())
    col_idx = torch.empty(nnz, dtype=torch.int32, device=w.device)
    blocks = torch.empty((nnz, bh, bw), dtype=torch.float32, device=w.device)
    if nnz:
        check(_lib.lib().accel_bsr_gather_f32(_ptr(w), rows, cols, w.stride(0), bh, bw, _ptr(slot), nbr, nbc,
                                              _ptr(col_idx), _ptr(blocks), _stream()))
    return row_ptr, col_idx, blocks


def row_absmax_f32(w: torch.Tensor) -> torch.Tensor:
    out = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
    check(_lib.lib().accel_row_absmax_f32(_ptr(w), w.shape[0], w.shape[1], w.stride(0), _ptr(out), _stream()))
    return out


def symmetric_scales_f32(absmax: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(absmax)
    check(_lib.lib().accel_symmetric_scales_f32(_ptr(absmax), absmax.numel(), _ptr(out), _stream()))
    return out


def quantize_rows_f32(w: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    q = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    check(_lib.lib().accel_quantize_rows_f32(_ptr(w), w.shape[0], w.shape[1], w.stride(0), _ptr(scales), _ptr(q), _stream()))
    return q


# ---------------------------------------------------------------------------------------- epilogue pieces / pools
def requant_i32_i8(acc: torch.Tensor, chan_scale: torch.Tensor, chan_dim: int, bias=None, relu=False,
                   sat_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    acc = acc.contiguous()
    n_chan = acc.shape[chan_dim]
    n_outer = int(np.prod(acc.shape[:chan_dim])) if chan_dim else 1
    n_inner = int(np.prod(acc.shape[chan_dim + 1:])) if chan_dim + 1 < acc.dim() else 1
    out = torch.empty(acc.shape, dtype=torch.int8, device=acc.device)
    check(_lib.lib().accel_requant_i32_i8(_ptr(acc), _ptr(out), n_outer, n_chan, n_inner, _ptr(chan_scale), _ptr(bias),
                                          int(relu), _ptr(sat_count), _stream()))
    return out


def add_residual_i8(a: torch.Tensor, b: torch.Tensor, s_main: float, s_res: float, s_out: float) -> torch.Tensor:
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    check(_lib.lib().accel_add_residual_i8(_ptr(a), _ptr(b), _ptr(out), a.numel(), s_main, s_res, s_out, _stream()))
    return out


def subsample2_int8(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``x[:, :, ::2, ::2]`` of an int8 NCHW tensor (dense or padded rows) into a padded-row tensor: the input a 1x1 / stride 2
    convolution really reads, so that it can run as a stride-1 pointwise convolution (csrc/conv_ws.cuh)."""
    if x.dtype != torch.int8 or x.dim() != 4 or not x.is_cuda:
        raise AcceleratorError(_lib.INVALID_CONFIG, "Activations must be a 4-D INT8 CUDA tensor (NCHW)")
    pitch = _row_pitch(x)
    if pitch is None:
        x = x.contiguous()
        pitch = x.shape[3]
    B, Cn, H, W = x.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    if out is None:
        out = alloc_padded((B, Cn, Ho, Wo))
    op = _row_pitch(out)
    if op is None or tuple(out.shape) != (B, Cn, Ho, Wo) or out.dtype != torch.int8:
        raise AcceleratorError(_lib.INVALID_CONFIG, "output must be int8 NCHW [B, C, ceil(H/2), ceil(W/2)], dense or with padded rows")
    check(_lib.lib().accel_subsample2_i8(_ptr(x), B * Cn, H, W, pitch, _ptr(out), op, _stream()))
    return out


def relu_int8(x: torch.Tensor) -> torch.Tensor:
    """relu_int8 (golden_models.cpp:278-283), in place on a contiguous CUDA int8 tensor; returns it."""
    _check_inplace(x, torch.int8)
    check(_lib.lib().accel_relu_i8(_ptr(x), x.numel(), _stream()))
    return x


def relu6_int8(x: torch.Tensor, scale: float) -> torch.Tensor:
    """relu6_int8 (golden_models.cpp:323-330): clamp to [0, int8(6.0f / scale)], in place; returns the tensor."""
    _check_inplace(x, torch.int8)
    check(_lib.lib().accel_relu6_i8(_ptr(x), x.numel(), float(scale), _stream()))
    return x


def relu_int32(x: torch.Tensor) -> torch.Tensor:
    """relu_int32 (golden_models.cpp:298-303), in place; returns the tensor."""
    _check_inplace(x, torch.int32)
    check(_lib.lib().accel_relu_i32(_ptr(x), x.numel(), _stream()))
    return x


def _check_inplace(x, dtype) -> None:
    if not _is_cuda(x) or x.dtype != dtype or not x.is_contiguous():
        raise AcceleratorError(_lib.INVALID_CONFIG, f"expected a contiguous CUDA {dtype} tensor")


def maxpool_i8(x: torch.Tensor, pool: int, stride: int, pad: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    in_pitch = _row_pitch(x) if x.dim() == 4 else None
    if in_pitch is None:
        x = x.contiguous()
        in_pitch = x.shape[-1]
    H, W = x.shape[-2:]
    Ho, Wo = (H + 2 * pad - pool) // stride + 1, (W + 2 * pad - pool) // stride + 1
    if out is None:
        out = torch.empty(x.shape[:-2] + (Ho, Wo), dtype=torch.int8, device=x.device)
    out_pitch = _row_pitch(out) if out.dim() == 4 else Wo
    if out_pitch is None:
        raise AcceleratorError(_lib.INVALID_CONFIG, "output must be dense or have padded rows")
    n_planes = int(np.prod(x.shape[:-2]))
    check(_lib.lib().accel_maxpool_i8(_ptr(x), _ptr(out), n_planes, H, W, pool, stride, pad, in_pitch, out_pitch, _stream()))
    return out


def avgpool_i8(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    in_pitch = _row_pitch(x) if x.dim() == 4 else None
    if in_pitch is None:
        x = x.contiguous()
        in_pitch = x.shape[-1]
    H, W = x.shape[-2:]
    if out is None:
        out = torch.empty(x.shape[:-2], dtype=torch.int8, device=x.device)
    check(_lib.lib().accel_avgpool_i8(_ptr(x), _ptr(out), int(np.prod(x.shape[:-2])), H, W, in_pitch, _stream()))
    return out
