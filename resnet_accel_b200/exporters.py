"""BSR packing format of the reference, produced on the GPU.

Same names, argument meaning, return dictionaries and on-disk layouts as
``sw/training/export_bsr_14x14.py``, ``sw/training/export_bsr.py``, ``sw/exporters/*`` and the
quantisers of ``sw/INT8 quantization/quantize.py`` - but the block statistics, the prefix scan,
the block gather and the per-row quantisation run in CUDA kernels (csrc/simple_kernels.cuh)
instead of Python double loops.  Results are returned as NumPy arrays (as the reference does)
unless ``device=True`` is passed, in which case the dictionary holds CUDA tensors ready for
``ops.BsrPlan`` without a host round trip.

Only host bookkeeping (dict assembly, JSON, file writes, the legacy ``np.random`` mask stream)
is Python.  There is no CPU fallback for the arithmetic.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import ops

BLOCK_SIZE = 14          # export_bsr_14x14.py:48
BLOCK_H = BLOCK_W = BLOCK_SIZE
BLOCK_ELEMENTS = BLOCK_H * BLOCK_W


def _dev(a, dtype) -> torch.Tensor:
    return ops.to_device(a, dtype)


def _as_2d(w):
    if hasattr(w, "ndim") and w.ndim == 4:           # conv OIHW -> [out, in*kh*kw] (export_bsr_14x14.py:554-558)
        return w.reshape(w.shape[0], -1)
    return w


def _result(data, indices, indptr, shape, bh, bw, device: bool) -> Dict:
    rows, cols = int(shape[0]), int(shape[1])
    padded = (-(-rows // bh) * bh, -(-cols // bw) * bw)
    nbr, nbc = padded[0] // bh, padded[1] // bw
    nnz = int(indices.numel())
    total = nbr * nbc
    density = nnz / total if total > 0 else 0.0
    if not device:
        data, indices, indptr = data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()
    return {
        "data": data, "indices": indices, "indptr": indptr,
        "shape": (rows, cols), "padded_shape": padded, "blocksize": (bh, bw),
        "num_blocks": nnz, "num_block_rows": nbr, "num_block_cols": nbc,
        "density": density, "sparsity_pct": (1.0 - density) * 100.0,
    }


# ------------------------------------------------------------------------------------------ packers
def build_bsr_14x14_int8_direct(weight_int8, threshold: float = 1e-10, device: bool = False) -> Dict:
    """export_bsr_14x14.py:406-484: keep a block iff sum(|block|) > threshold (int32 sum)."""
    w = _dev(_as_2d(weight_int8), torch.int8)
    l1 = ops.block_l1_i8(w, BLOCK_SIZE)
    keep = l1.to(torch.float64) > float(threshold)
    rp, ci, blocks = ops.pack_bsr_i8(w, keep, BLOCK_SIZE)
    return _result(blocks, ci, rp, w.shape, BLOCK_H, BLOCK_W, device)


def build_bsr_14x14(weight, threshold: float = 1e-10, quantize: bool = False, scale=None,
                    device: bool = False) -> Dict:
    """export_bsr_14x14.py:84-235: FP32 blocks kept iff L2 norm > threshold; optional per-row INT8
    quantisation with ``scale[global_row]`` (``scale[0]`` for rows beyond the scales, :181-195)."""
    w = _dev(_as_2d(weight), torch.float32)
    rows, cols = w.shape
    keep = ops.block_l2_f32(w, BLOCK_H, BLOCK_W) > float(threshold)
    if not quantize:
        rp, ci, blocks = ops.pack_bsr_f32(w, keep, BLOCK_H, BLOCK_W)
        return _result(blocks, ci, rp, w.shape, BLOCK_H, BLOCK_W, device)
    if scale is None:
        raise ValueError("scale required for quantization")
    s = _dev(scale, torch.float32).reshape(-1)
    if s.numel() < rows:                                 # rows past the scales use scale[0] (:189-190)
        fill = s[:1] if s.numel() else torch.ones(1, dtype=torch.float32, device=w.device)
        s = torch.cat([s, fill.expand(rows - s.numel())])
    q = ops.quantize_rows_f32(w, s[:rows].contiguous())
    rp, ci, blocks = ops.pack_bsr_i8(q, keep, BLOCK_SIZE)
    return _result(blocks, ci, rp, w.shape, BLOCK_H, BLOCK_W, device)


def build_bsr_from_dense(weight, block_h: int, block_w: int, threshold: float = 1e-10, device: bool = False) -> Dict:
    """export_bsr.py:76-153 (generic block shape): keep iff Frobenius norm > threshold; ``data`` keeps
    the input's element type (float32 or int8)."""
    w2 = _as_2d(weight)
    is_int8 = (isinstance(w2, np.ndarray) and w2.dtype == np.int8) or (isinstance(w2, torch.Tensor) and w2.dtype == torch.int8)
    if is_int8:
        w = _dev(w2, torch.int8)
        keep = ops.block_l2_f32(w.to(torch.float32), block_h, block_w) > float(threshold)
        if block_h != block_w:
            raise ValueError("int8 packing supports square blocks only")
        rp, ci, blocks = ops.pack_bsr_i8(w, keep, block_h)
    else:
        w = _dev(w2, torch.float32)
        keep = ops.block_l2_f32(w, block_h, block_w) > float(threshold)
        rp, ci, blocks = ops.pack_bsr_f32(w, keep, block_h, block_w)
    return _result(blocks, ci, rp, w.shape, block_h, block_w, device)


# ------------------------------------------------------------------------------------------ quantisers
def quantize_symmetric_per_channel(x, axis: int = 0, device: bool = False) -> Tuple:
    """quantize.py:71-98: scales = max(maxabs/127, 1e-12) (float32), q = clip(rint(x/scale))."""
    if axis != 0:
        raise ValueError("only axis=0 (output channels) is supported")
    shape = tuple(x.shape)
    w = _dev(_as_2d(x) if len(shape) != 2 else x, torch.float32).reshape(shape[0], -1)
    absmax = ops.row_absmax_f32(w)
    scales = ops.symmetric_scales_f32(absmax)            # IEEE float32 divide by 127, floor 1e-12 (quantize.py:86)
    q = ops.quantize_rows_f32(w, scales).reshape(shape)
    if device:
        return q, scales
    return q.cpu().numpy(), scales.cpu().numpy()


def quantize_symmetric_per_tensor(x) -> Tuple[np.ndarray, float]:
    """quantize.py:55-68 (scale is a Python float = float64 arithmetic on the max)."""
    w = _dev(x, torch.float32)
    shape = tuple(w.shape)
    w2 = w.reshape(1, -1)
    maxabs = float(ops.row_absmax_f32(w2)[0].item())
    scale = max(maxabs / 127.0, 1e-12)
    # x / scale is evaluated in float32 against the float64 scale rounded to float32 by NumPy's weak-scalar rule
    s = torch.full((1,), np.float32(scale), dtype=torch.float32, device=w.device)
    return ops.quantize_rows_f32(w2, s).reshape(shape).cpu().numpy(), scale


# ------------------------------------------------------------------------------------------ masks / pruning
def create_sparse_mask(shape: Tuple[int, int], sparsity_pct: float, block_size: int = 8, seed: int = 42) -> np.ndarray:
    """sw/exporters/export_transformer.py:19-60.  Host logic: it must replay NumPy's legacy global
    RNG stream (``np.random.seed`` + ``np.random.choice``) to give the reference's masks."""
    np.random.seed(seed)
    rows, cols = shape
    nbr, nbc = -(-rows // block_size), -(-cols // block_size)
    total = nbr * nbc
    zero = np.random.choice(total, size=int(total * sparsity_pct / 100.0), replace=False)
    keep = np.ones(total, dtype=bool)
    keep[zero] = False
    full = np.kron(keep.reshape(nbr, nbc), np.ones((block_size, block_size), dtype=bool))
    return full[:rows, :cols].astype(np.float32)


def compute_block_norms(weight, block_h: int, block_w: int):
    """blocksparse_train.py:93-138 on the GPU: (norms [nbh, nbw] float32 CUDA tensor, (nbh, nbw), original shape)."""
    shape = tuple(weight.shape)
    w = _dev(_as_2d(weight), torch.float32)
    norms = ops.block_l2_f32(w, block_h, block_w)
    return norms, tuple(norms.shape), shape


def prune_blocks_global(norms_per_layer, target_sparsity: float, min_keep):
    """Selection rule of blocksparse_train.py:141-239 without the per-block host syncs: one stable
    device sort over all layers' norms, then the floor-respecting walk over the sorted order.
    Returns a list of boolean keep masks (CUDA tensors), one per layer."""
    dev = norms_per_layer[0].device
    flat = torch.cat([n.reshape(-1) for n in norms_per_layer])
    layer_of = torch.cat([torch.full((n.numel(),), i, dtype=torch.int64, device=dev) for i, n in enumerate(norms_per_layer)])
    order = torch.sort(flat, stable=True).indices
    budget = int(flat.numel() * target_sparsity)
    totals = [int(n.numel()) for n in norms_per_layer]
    floors = [int(t * mk) for t, mk in zip(totals, min_keep)]
    max_prune = [t - f for t, f in zip(totals, floors)]
    # rank of each candidate inside its own layer along the sorted order; a layer stops giving once
    # it has given max_prune blocks - identical to the reference's sequential "skip if at floor"
    lay_sorted = layer_of[order]
    prune_sorted = torch.zeros_like(lay_sorted, dtype=torch.bool)
    eligible = torch.zeros_like(lay_sorted, dtype=torch.bool)
    for i, mp in enumerate(max_prune):
        sel = lay_sorted == i
        rank = torch.cumsum(sel.to(torch.int64), 0) - 1
        eligible |= sel & (rank < mp)
    # NB: the reference appends before it tests `len(to_prune) >= budget`, so a budget of 0 still prunes one block
    take = torch.cumsum(eligible.to(torch.int64), 0) <= max(budget, 1)
    prune_sorted = eligible & take
    keep_flat = torch.ones_like(flat, dtype=torch.bool)
    keep_flat[order[prune_sorted]] = False
    out, o = [], 0
    for n in norms_per_layer:
        out.append(keep_flat[o:o + n.numel()].reshape(n.shape))
        o += n.numel()
    return out


# ------------------------------------------------------------------------------------------ file formats
def save_bsr_binary_int8(bsr_data: Dict, filepath: str):
    """export_bsr_14x14.py:241-272: raw int8 blocks, 196 B each, row-major inside the block, no header."""
    data = bsr_data["data"]
    data = data.cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
    if data.dtype != np.int8:
        raise ValueError(f"Expected INT8 data, got {data.dtype}")
    with open(filepath, "wb") as f:
        f.write(np.ascontiguousarray(data).tobytes())


def save_bsr_metadata(bsr_data: Dict, filepath: str, layer_name: str):
    """export_bsr_14x14.py:274-317 (same keys)."""
    indptr = np.asarray(bsr_data["indptr"].cpu() if isinstance(bsr_data["indptr"], torch.Tensor) else bsr_data["indptr"])
    indices = np.asarray(bsr_data["indices"].cpu() if isinstance(bsr_data["indices"], torch.Tensor) else bsr_data["indices"])
    meta = {
        "layer_name": layer_name,
        "shape": list(bsr_data["shape"]),
        "padded_shape": list(bsr_data["padded_shape"]),
        "blocksize": list(bsr_data["blocksize"]),
        "num_blocks": int(bsr_data["num_blocks"]),
        "num_block_rows": int(bsr_data["num_block_rows"]),
        "num_block_cols": int(bsr_data["num_block_cols"]),
        "density": float(bsr_data["density"]),
        "sparsity_pct": float(bsr_data["sparsity_pct"]),
        "row_ptr": indptr.tolist(),
        "col_idx": indices.tolist(),
        "tiles_per_row": np.diff(indptr).astype(int).tolist(),
        "max_tiles_per_row": int(np.max(np.diff(indptr))) if len(indptr) > 1 else 0,
        "bytes_per_block": BLOCK_ELEMENTS,
        "total_weight_bytes": int(bsr_data["num_blocks"] * BLOCK_ELEMENTS),
    }
    with open(filepath, "w") as f:
        json.dump(meta, f, indent=2)


def export_int8_layer_14x14(name: str, weight_int8, output_dir: str) -> Dict:
    """export_bsr_14x14.py:487-523: weights.bsr + row_ptr.npy + col_idx.npy + weights.meta.json."""
    bsr = build_bsr_14x14_int8_direct(weight_int8)
    layer_dir = os.path.join(output_dir, name)
    os.makedirs(layer_dir, exist_ok=True)
    save_bsr_binary_int8(bsr, os.path.join(layer_dir, "weights.bsr"))
    np.save(os.path.join(layer_dir, "row_ptr.npy"), bsr["indptr"])
    np.save(os.path.join(layer_dir, "col_idx.npy"), bsr["indices"])
    save_bsr_metadata(bsr, os.path.join(layer_dir, "weights.meta.json"), name)
    return bsr


def export_layer_14x14(name: str, weight, output_dir: str, scales=None) -> Dict:
    """export_bsr_14x14.py:323-400: FP32 layer -> (optionally quantised) 14x14 BSR files."""
    w2 = _as_2d(weight)
    bsr = build_bsr_14x14(w2, quantize=scales is not None, scale=scales)
    layer_dir = os.path.join(output_dir, name)
    os.makedirs(layer_dir, exist_ok=True)
    if scales is not None:
        save_bsr_binary_int8(bsr, os.path.join(layer_dir, "weights.bsr"))
    else:
        with open(os.path.join(layer_dir, "weights.bsr"), "wb") as f:
            f.write(np.ascontiguousarray(bsr["data"], dtype=np.float32).tobytes())
    np.save(os.path.join(layer_dir, "row_ptr.npy"), bsr["indptr"])
    np.save(os.path.join(layer_dir, "col_idx.npy"), bsr["indices"])
    save_bsr_metadata(bsr, os.path.join(layer_dir, "weights.meta.json"), name)
    return bsr


def export_from_int8_dir(int8_dir: str, output_dir: str) -> Dict:
    """export_bsr_14x14.py:526-602: every ``<layer>_weight_int8.npy`` -> 14x14 BSR + model_summary.json."""
    os.makedirs(output_dir, exist_ok=True)
    stats, total, nonzero = [], 0, 0
    for name in ("conv1", "conv2", "fc1", "fc2"):
        path = os.path.join(int8_dir, f"{name}_weight_int8.npy")
        if not os.path.exists(path):
            continue
        w = np.load(path)
        bsr = export_int8_layer_14x14(name, _as_2d(w), output_dir)
        grid = bsr["num_block_rows"] * bsr["num_block_cols"]
        stats.append({"name": name, "original_shape": list(bsr["shape"]), "padded_shape": list(bsr["padded_shape"]),
                      "blocksize": [BLOCK_H, BLOCK_W], "num_blocks": bsr["num_blocks"], "total_blocks": grid,
                      "density": bsr["density"], "sparsity_pct": bsr["sparsity_pct"]})
        total += grid
        nonzero += bsr["num_blocks"]
    summary = {"model": "MNIST CNN INT8 (14×14 BSR)", "hardware_block_size": BLOCK_SIZE, "source": "int8_quantized",
               "total_blocks": total, "nonzero_blocks": nonzero,
               "overall_density": nonzero / total if total else 0.0,
               "overall_sparsity_pct": (1.0 - nonzero / total) * 100 if total else 0.0, "layers": stats}
    with open(os.path.join(output_dir, "model_summary.json"), "w") as f:
        json.dump(summary, f, indent=2)
    return summary


def serialize_for_hardware(bsr: Dict) -> bytes:
    """bsr_packer.hpp:489-525 / export_resnet18_bsr.py:199-247: u32 nnz, nbr, nbc; u16 row_ptr, col_idx; int8 data."""
    g = lambda k: (bsr[k].cpu().numpy() if isinstance(bsr[k], torch.Tensor) else np.asarray(bsr[k]))
    if bsr["num_blocks"] >= 65536:
        raise ValueError("u16 row_ptr cannot index >= 65536 blocks")
    hdr = np.array([bsr["num_blocks"], bsr["num_block_rows"], bsr["num_block_cols"]], dtype="<u4").tobytes()
    return hdr + g("indptr").astype("<u2").tobytes() + g("indices").astype("<u2").tobytes() + g("data").astype(np.int8).tobytes()
