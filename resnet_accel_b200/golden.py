"""The reference's golden-model layer calls, computed on the B200.

Same names and positional signatures as ``sw/golden/golden_fc1_test.py`` and
``sw/golden/gemm_bsr_int8.py``; NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out.
"""
from __future__ import annotations

import json
import os
from typing import Dict

import numpy as np
import torch

from . import ops
from ._lib import AcceleratorError, INVALID_CONFIG


def load_bsr_layer(layer_dir: str) -> Dict:
    """golden_fc1_test.py:16-46: row_ptr.npy, col_idx.npy, raw int8 weights.bsr, weights.meta.json."""
    with open(os.path.join(layer_dir, "weights.meta.json"), "r") as f:
        meta = json.load(f)
    block_h, block_w = meta["blocksize"]
    num_blocks = meta["num_blocks"]
    with open(os.path.join(layer_dir, "weights.bsr"), "rb") as f:
        flat = np.frombuffer(f.read(), dtype=np.int8)
    return {
        "row_ptr": np.load(os.path.join(layer_dir, "row_ptr.npy")),
        "col_idx": np.load(os.path.join(layer_dir, "col_idx.npy")),
        "weights": flat.reshape(num_blocks, block_h, block_w),
        "block_h": block_h,
        "block_w": block_w,
        "num_blocks": num_blocks,
        "shape": meta["padded_shape"],
        "original_shape": meta["shape"],
    }


def _plan_for(bsr_layer: Dict) -> ops.BsrPlan:
    plan = bsr_layer.get("_plan")
    if plan is None:
        nbc = None
        if "shape" in bsr_layer and bsr_layer.get("block_w"):
            nbc = int(bsr_layer["shape"][1]) // int(bsr_layer["block_w"])
        plan = ops.BsrPlan(bsr_layer["row_ptr"], bsr_layer["col_idx"], bsr_layer["weights"], n_block_cols=nbc)
        bsr_layer["_plan"] = plan          # weights stay resident across calls (weight-stationary)
    return plan


def gemm_bsr_int8_golden(activations, bsr_layer: Dict):
    """golden_fc1_test.py:49-108: INT8 activations [M, K] x BSR weights -> INT32 [M, n_block_rows*block_h].

    14x14 blocks run on the tcgen05 kernel; any other block size on the CUDA-core generic kernel."""
    bh, bw = int(bsr_layer["block_h"]), int(bsr_layer["block_w"])
    as_numpy = not isinstance(activations, torch.Tensor)
    x = ops.to_device(activations, torch.int8)
    if x.dim() != 2:
        raise AcceleratorError(INVALID_CONFIG, "activations must be [M, K]")
    nbr = len(bsr_layer["row_ptr"]) - 1
    if bh == 14 and bw == 14:
        plan = _plan_for(bsr_layer)
        if -(-x.shape[1] // 14) > plan.n_block_cols:      # activations wider than the weight grid: extra K is unused
            x = x[:, :plan.n_block_cols * 14]
        out = plan.gemm(x, "i32")
    else:
        out = ops.bsr_gemm_generic(x, bsr_layer["row_ptr"], bsr_layer["col_idx"], bsr_layer["weights"], nbr * bh)
    return out.cpu().numpy() if as_numpy else out


def _is_f64_scalar(v) -> bool:
    """NumPy promotion (NEP 50): Python floats are weak (the float32 array keeps its dtype), np.float64 scalars and 0-d
    float64 arrays are strong (the product is carried in float64)."""
    if isinstance(v, np.ndarray):
        return v.dtype == np.float64
    return isinstance(v, np.floating) and np.dtype(type(v)) == np.float64


def gemm_bsr_int8(A_int8, bsr_B: Dict, scale_A, scales_B):
    """sw/golden/gemm_bsr_int8.py:16-104 on the GPU: the reference's float32 'golden' with per-row scales.

    Its arithmetic is not ``A @ B``: every stored block is re-quantised with the scales of its rows, the
    INT32 tile ``A[:, block-row span] @ block^T`` is added to the output once per local row with that row's scale, in
    float32 (or float64, when the caller's scales are float64) and in (block-row, block, local-row) order.  The reference's
    own tests only pin its output shape; this replay is bit-exact against the reference function itself
    (tests/golden/golden_fp32_cases.npz).  A compatibility path on CUDA cores - the numerical oracle is
    :func:`gemm_bsr_int8_golden`."""
    import ctypes as C
    from . import _lib
    as_numpy = not isinstance(A_int8, torch.Tensor)
    dev = ops._require_cuda()
    data = bsr_B["data"]
    data_np = data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    scales_np = scales_B.detach().cpu().numpy() if isinstance(scales_B, torch.Tensor) else np.asarray(scales_B)
    if scales_np.ndim != 1 or scales_np.size == 0:
        raise ValueError("scales_B must be a non-empty 1-D array")
    K, N = (int(v) for v in bsr_B["shape"])
    bh, bw = (int(v) for v in bsr_B["blocksize"])
    if bh != bw:
        raise ValueError(f"matmul: A[:, block rows] @ block.T needs square blocks, got {bh}x{bw}")     # NumPy raises here too
    if K % bh:
        raise ValueError("matmul: the last block-row is partial (K is not a multiple of the block size)")
    x = ops.to_device(A_int8, torch.int8, dev)
    if x.dim() != 2 or x.shape[1] < K:
        raise ValueError("A_int8 must be [M, K]")
    indptr = np.ascontiguousarray(np.asarray(ops._host(bsr_B["indptr"])), dtype=np.int32)
    indices = np.ascontiguousarray(np.asarray(ops._host(bsr_B["indices"])), dtype=np.int32)
    nbr, nnz = len(indptr) - 1, int(indices.size)
    s64 = scales_np.dtype == np.float64
    div64 = s64 or data_np.dtype == np.float64
    if scales_np.dtype not in (np.float32, np.float64):
        scales_np = scales_np.astype(np.float32)
    fdt = np.float64 if div64 else np.float32
    data_d = ops.to_device(np.ascontiguousarray(data_np.reshape(-1), dtype=fdt), torch.float64 if div64 else torch.float32, dev)
    sdiv_d = ops.to_device(np.ascontiguousarray(scales_np, dtype=fdt), torch.float64 if div64 else torch.float32, dev)
    s_d = ops.to_device(np.ascontiguousarray(scales_np), torch.float64 if s64 else torch.float32, dev)
    blk_row = np.repeat(np.arange(nbr, dtype=np.int32), np.diff(indptr))
    ip_d, ix_d, br_d = (ops.to_device(a, torch.int32, dev) for a in (indptr, indices if nnz else np.zeros(1, np.int32),
                                                                    blk_row if nnz else np.zeros(1, np.int32)))
    q = torch.empty(max(nnz * bh * bw, 1), dtype=torch.int8, device=dev)
    M = int(x.shape[0])
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    a64 = _is_f64_scalar(scale_A)
    _lib.check(_lib.lib().accel_gemm_bsr_int8_fp32(
        ops._ptr(x), M, x.stride(0) if M else K, ops._ptr(ip_d), ops._ptr(ix_d), ops._ptr(data_d), ops._ptr(sdiv_d), int(div64),
        ops._ptr(br_d), nnz, nbr, bh, K, N, C.c_double(float(scale_A)), int(a64), ops._ptr(s_d), int(s64), int(scales_np.size),
        ops._ptr(q), ops._ptr(out), N, ops._stream()))
    return out.cpu().numpy() if as_numpy else out
