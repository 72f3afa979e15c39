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


def gemm_bsr_int8(A_int8, bsr_B: Dict, scale_A, scales_B):
    """sw/golden/gemm_bsr_int8.py:16-104 is a float32 'golden' whose arithmetic is order-dependent and
    unpinned by the reference's tests.  This build offers the INT32 path
    (:func:`gemm_bsr_int8_golden`) and the fused per-channel de-quantisation (``BsrPlan.gemm(..., "f32")``);
    the quirky replay is deliberately not provided on the device."""
    raise AcceleratorError(INVALID_CONFIG,
                           "gemm_bsr_int8 (FP32 compat golden) is not part of the device path; use "
                           "gemm_bsr_int8_golden or BsrPlan.gemm(out_kind='f32')")
