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


def deserialize_from_hardware(buffer, block: int = BLOCK_SIZE) -> Dict:
    """bsr_packer.hpp:530-575: the reader of :func:`serialize_for_hardware` (little-endian ``u32 nnz, nbr, nbc``, ``u16
    row_ptr[nbr+1]``, ``u16 col_idx[nnz]``, ``int8 data[nnz * block * block]``).  Raises ``ValueError`` where the C++ throws
    (``"Buffer too small for BSR header"`` / ``"Buffer size mismatch: expected N, got M"``)."""
    buf = bytes(buffer)
    if len(buf) < 12:
        raise ValueError("Buffer too small for BSR header")
    nnz, nbr, nbc = (int(v) for v in np.frombuffer(buf, dtype="<u4", count=3))
    be = block * block
    expected = 12 + (nbr + 1) * 2 + nnz * 2 + nnz * be
    if len(buf) < expected:
        raise ValueError(f"Buffer size mismatch: expected {expected}, got {len(buf)}")
    off = 12
    indptr = np.frombuffer(buf, dtype="<u2", count=nbr + 1, offset=off).astype(np.int32)
    off += (nbr + 1) * 2
    indices = np.frombuffer(buf, dtype="<u2", count=nnz, offset=off).astype(np.int32)
    off += nnz * 2
    data = np.frombuffer(buf, dtype=np.int8, count=nnz * be, offset=off).reshape(nnz, block, block).copy()
    total = nbr * nbc
    density = nnz / total if total else 0.0
    return {"data": data, "indices": indices, "indptr": indptr, "blocksize": (block, block), "num_blocks": nnz,
            "num_block_rows": nbr, "num_block_cols": nbc, "padded_shape": (nbr * block, nbc * block),
            "shape": (nbr * block, nbc * block), "density": density, "sparsity_pct": (1.0 - density) * 100.0}


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def quantize_blocks_by_block_row(bsr_data: Dict, scales) -> np.ndarray:
    """The quantisation inside ``save_bsr_binary_int8(bsr, scales, path)`` of sw/training/export_bsr.py:177-202: every stored
    block is divided by ONE scale, ``scales[block_row]`` (the index of its block-row, not of its channels - the quirk the
    shipped ``data/fixtures`` depend on; ``scales[0]`` past the end, a 0-d ``scales`` for everything), then
    ``clip(rint(.), -128, 127)``.  float32 blocks with float32 scales run through the GPU row quantiser
    (``accel_quantize_rows_f32``: one "row" per block); as soon as either side is float64 NumPy divides in float64, and so
    does this (device float64 ops).  Returns int8 [nnz, bh, bw]."""
    data = _np(bsr_data["data"])
    indptr = _np(bsr_data["indptr"]).astype(np.int64)
    nnz = int(data.shape[0])
    bh, bw = (int(v) for v in bsr_data["blocksize"])
    if nnz == 0:
        return np.zeros((0, bh, bw), np.int8)
    sc = _np(scales)
    row_of_block = np.searchsorted(indptr, np.arange(nnz) + 1) - 1                   # export_bsr.py:188
    if sc.ndim > 0:
        per_block = np.where(row_of_block < len(sc), sc[np.minimum(row_of_block, len(sc) - 1)], sc[0])
    else:
        per_block = np.full(nnz, float(sc))                                          # float(scales): a Python float, weak
    f64 = data.dtype == np.float64 or (sc.ndim > 0 and sc.dtype == np.float64)
    dev = ops._require_cuda()
    if not f64:
        d = ops.to_device(np.ascontiguousarray(data.reshape(nnz, bh * bw), dtype=np.float32), torch.float32, dev)
        s = ops.to_device(np.ascontiguousarray(per_block, dtype=np.float32), torch.float32, dev)
        return ops.quantize_rows_f32(d, s).reshape(nnz, bh, bw).cpu().numpy()
    d = torch.from_numpy(np.ascontiguousarray(data.reshape(nnz, bh * bw), dtype=np.float64)).to(dev)
    s = torch.from_numpy(np.ascontiguousarray(per_block, dtype=np.float64)).to(dev)
    q = torch.clamp(torch.round(d / s[:, None]), -128, 127).to(torch.int8)          # IEEE divide, round half to even
    return q.reshape(nnz, bh, bw).cpu().numpy()


def save_bsr_binary(bsr_data: Dict, filepath: str) -> None:
    """export_bsr.py:158-175: the blocks as float32, row-major inside the block, no header."""
    with open(filepath, "wb") as f:
        f.write(np.ascontiguousarray(_np(bsr_data["data"]), dtype=np.float32).tobytes())


def save_bsr_binary_int8_generic(bsr_data: Dict, scales, filepath: str) -> None:
    """export_bsr.py:177-202 (the 3-argument ``save_bsr_binary_int8`` of the generic exporter): raw int8 blocks quantised
    with the block-row-indexed scale (:func:`quantize_blocks_by_block_row`), no header."""
    with open(filepath, "wb") as f:
        f.write(quantize_blocks_by_block_row(bsr_data, scales).tobytes())


def save_bsr_metadata_generic(bsr_data: Dict, filepath: str, layer_name: str = "") -> None:
    """export_bsr.py:205-241 (same keys, incl. ``avg_tiles_per_row``; ``sparsity_pct`` recomputed from the density)."""
    indptr, indices = _np(bsr_data["indptr"]), _np(bsr_data["indices"])
    nbr = int(bsr_data["num_block_rows"])
    meta = {
        "layer_name": layer_name,
        "shape": list(bsr_data["shape"]),
        "padded_shape": list(bsr_data.get("padded_shape", bsr_data["shape"])),
        "blocksize": list(bsr_data["blocksize"]),
        "num_blocks": int(bsr_data["num_blocks"]),
        "num_block_rows": nbr,
        "num_block_cols": int(bsr_data["num_block_cols"]),
        "density": float(bsr_data["density"]),
        "sparsity_pct": float((1.0 - bsr_data["density"]) * 100),
        "row_ptr": indptr.tolist(),
        "col_idx": indices.tolist(),
        "tiles_per_row": [int(indptr[i + 1] - indptr[i]) for i in range(nbr)],
        "max_tiles_per_row": int(np.max(np.diff(indptr))) if len(indptr) > 1 else 0,
        "avg_tiles_per_row": float(bsr_data["num_blocks"] / nbr) if nbr > 0 else 0.0,
    }
    with open(filepath, "w") as f:
        json.dump(meta, f, indent=2)


def quantize_blocks_per_channel(bsr: Dict, scales) -> np.ndarray:
    """The quantisation inside ``save_bsr_binary_int8`` of sw/training/export_resnet18_bsr.py:199-247: row ``r`` of a block in
    block-row ``br`` is channel ``br * bh + r`` and is divided by that channel's scale; rows whose channel lies beyond
    ``len(scales)`` stay zero (:233-239).  GPU row quantiser over the [nnz * bh, bw] view.  Returns int8 [nnz, bh, bw]."""
    data = _np(bsr["data"])
    indptr = _np(bsr["indptr"]).astype(np.int64)
    nnz = int(data.shape[0])
    bh, bw = (int(v) for v in bsr["blocksize"])
    if nnz == 0:
        return np.zeros((0, bh, bw), np.int8)
    sc = _np(scales).reshape(-1)
    block_row = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    chan = (block_row[:, None] * bh + np.arange(bh)[None, :]).reshape(-1)            # channel of every block row
    valid = chan < len(sc)
    row_scale = np.where(valid, sc[np.minimum(chan, len(sc) - 1)], 1.0)
    dev = ops._require_cuda()
    f64 = data.dtype == np.float64 or sc.dtype == np.float64
    if not f64:
        d = ops.to_device(np.ascontiguousarray(data.reshape(nnz * bh, bw), dtype=np.float32), torch.float32, dev)
        q = ops.quantize_rows_f32(d, ops.to_device(np.ascontiguousarray(row_scale, dtype=np.float32), torch.float32, dev))
    else:
        d = torch.from_numpy(np.ascontiguousarray(data.reshape(nnz * bh, bw), dtype=np.float64)).to(dev)
        s = torch.from_numpy(np.ascontiguousarray(row_scale, dtype=np.float64)).to(dev)
        q = torch.clamp(torch.round(d / s[:, None]), -128, 127).to(torch.int8)
    q = q.cpu().numpy().reshape(nnz * bh, bw)
    q[~valid] = 0
    return q.reshape(nnz, bh, bw)


def save_bsr_binary_int8_resnet(bsr: Dict, scales, filepath: str) -> None:
    """export_resnet18_bsr.py:199-247: the headered hardware blob (== ``serialize_for_hardware``, bsr_packer.hpp:489-525) with
    the blocks quantised per output channel on the way out.  ``struct.pack('<H', v)`` in the reference raises for
    ``v > 65535``; so does this."""
    indptr, indices = _np(bsr["indptr"]).astype(np.int64), _np(bsr["indices"]).astype(np.int64)
    if (indptr > 65535).any() or (indices > 65535).any() or (indptr < 0).any() or (indices < 0).any():
        raise ValueError("'H' format requires 0 <= number <= 65535")
    q = quantize_blocks_per_channel(bsr, scales)
    with open(filepath, "wb") as f:
        f.write(np.array([bsr["num_blocks"], bsr["num_block_rows"], bsr["num_block_cols"]], dtype="<u4").tobytes())
        f.write(indptr.astype("<u2").tobytes())
        f.write(indices.astype("<u2").tobytes())
        f.write(q.tobytes())


def save_bsr_metadata_resnet(bsr: Dict, filepath: str, layer_name: str = "") -> None:
    """export_resnet18_bsr.py:250-266."""
    meta = {"layer_name": layer_name, "original_shape": list(bsr["shape"]), "padded_shape": list(bsr["padded_shape"]),
            "blocksize": list(bsr["blocksize"]), "num_blocks": int(bsr["num_blocks"]),
            "num_block_rows": int(bsr["num_block_rows"]), "num_block_cols": int(bsr["num_block_cols"]),
            "sparsity_pct": float(bsr["sparsity_pct"]), "row_ptr": _np(bsr["indptr"]).tolist(),
            "col_idx": _np(bsr["indices"]).tolist()}
    with open(filepath, "w") as f:
        json.dump(meta, f, indent=2)


def get_resnet18_layer_config() -> Dict[str, Dict]:
    """export_resnet18_bsr.py:49-92: block shape and keep floor per ResNet-18 layer of the reference's ResNet exporter
    (convolutions 4x4, fc 14x14; floors 0.5 for the stem and the downsamples, 0.30 / 0.20 / 0.15 / 0.10 for the stages,
    0.05 for the fc).  The north-star path uses 14x14 everywhere (``export_bsr_14x14.py:661-673``); this table serves the
    ResNet file format and the pruning floors."""
    cfg = {"conv1": {"block_size": (4, 4), "min_keep": 0.50, "type": "conv"}}
    for stage, floor in ((1, 0.30), (2, 0.20), (3, 0.15), (4, 0.10)):
        for blk in (0, 1):
            for conv in ("conv1", "conv2"):
                cfg[f"layer{stage}.{blk}.{conv}"] = {"block_size": (4, 4), "min_keep": floor, "type": "conv"}
            if blk == 0 and stage > 1:
                cfg[f"layer{stage}.0.downsample.0"] = {"block_size": (4, 4), "min_keep": 0.50, "type": "conv"}
    cfg["fc"] = {"block_size": (14, 14), "min_keep": 0.05, "type": "linear"}
    return cfg


# ------------------------------------------------------------------------------------------ masks / pruning schedule
def create_block_sparse_mask(weight, block_size: Tuple[int, int], sparsity: float, seed: int = 42) -> torch.Tensor:
    """sw/training/train_resnet18.py:69-132.  Host logic for the RNG (it must replay ``torch.manual_seed(seed)`` +
    ``torch.randperm(total)`` on the CPU generator, as the reference does); the expansion of the block mask to element level
    is one ``repeat_interleave`` on the weight's device instead of the reference's Python loop over blocks.  Returns a mask of
    the weight's shape, dtype and device with 1.0 inside kept blocks."""
    torch.manual_seed(seed)
    w = weight if isinstance(weight, torch.Tensor) else torch.as_tensor(np.asarray(weight))
    shape = tuple(w.shape)
    w2 = w.reshape(shape[0], -1) if w.dim() == 4 else w
    rows, cols = w2.shape
    bh, bw = block_size
    nbr, nbc = -(-rows // bh), -(-cols // bw)
    total = nbr * nbc
    keep = torch.ones(total, dtype=torch.bool)
    keep[torch.randperm(total)[: int(total * sparsity)]] = False
    full = keep.reshape(nbr, nbc).to(w.device).repeat_interleave(bh, 0).repeat_interleave(bw, 1)[:rows, :cols]
    return full.to(w.dtype).reshape(shape)


def apply_masks(weights: Dict[str, torch.Tensor], masks: Dict[str, torch.Tensor]) -> None:
    """blocksparse_train.py:242-248 without the ``nn.Module`` walk: ``weight.mul_(mask)`` in place for every named weight."""
    with torch.no_grad():
        for name, w in weights.items():
            w.mul_(masks[name].to(w.dtype).reshape(w.shape))


def layer_block_cfg(name: str, weight) -> Tuple[Tuple[int, int], float]:
    """blocksparse_train.py:52-57: block shape and keep floor of the MNIST pruning flow - convolutions (4-D weights) 4x4 with
    at least 30 % of the blocks kept, linear layers 8x8 with at least 5 %.  Returns ``((block_h, block_w), min_keep)``."""
    if getattr(weight, "ndim", 2) == 4:
        return (4, 4), 0.30
    return (8, 8), 0.05


SPARSITY_SCHEDULE = (0.5, 0.7, 0.85)          # blocksparse_train.py:295, followed by the final target


def progressive_sparsity(weights: Dict[str, torch.Tensor], masks: Dict[str, torch.Tensor], target_sparsity: float = 0.9,
                         block_cfg=None, fine_tune=None) -> list:
    """blocksparse_train.py:282-321 on the GPU: 50 % -> 70 % -> 85 % -> target.  Each phase computes the block norms of the
    CURRENT (masked) weights (``accel_block_l2_f32``), selects globally with the per-layer floors
    (:func:`prune_blocks_global`), clears the pruned blocks in the persistent masks and applies them (:func:`apply_masks`).
    ``fine_tune(weights, masks, sparsity)`` stands where the reference calls ``train_with_group_lasso`` (the training loop
    itself is out of scope); it may update ``weights`` in place between phases.
    Returns one record per phase: ``{"sparsity", "blocks_pruned", "kept_per_layer"}``."""
    cfg = block_cfg or layer_block_cfg
    names = list(weights.keys())
    history = []
    for sparsity in tuple(SPARSITY_SCHEDULE) + (target_sparsity,):
        norms, shapes = [], []
        for n in names:
            (bh, bw), _ = cfg(n, weights[n])
            nm, _, _ = compute_block_norms(weights[n], bh, bw)
            norms.append(nm)
            shapes.append((bh, bw))
        keeps = prune_blocks_global(norms, sparsity, [cfg(n, weights[n])[1] for n in names])
        pruned = 0
        kept = {}
        for n, keep, (bh, bw) in zip(names, keeps, shapes):
            w = weights[n]
            rows, cols = w.shape[0], w.numel() // w.shape[0]
            full = keep.repeat_interleave(bh, 0).repeat_interleave(bw, 1)[:rows, :cols].reshape(w.shape)
            m = masks[n]
            masks[n] = (m.to(torch.bool).reshape(w.shape) & full.to(m.device)).to(m.dtype)
            pruned += int((~keep).sum().item())
            kept[n] = int(keep.sum().item())
        apply_masks(weights, masks)
        history.append({"sparsity": float(sparsity), "blocks_pruned": pruned, "kept_per_layer": kept})
        if fine_tune is not None:
            fine_tune(weights, masks, sparsity)
            apply_masks(weights, masks)
    return history


# ------------------------------------------------------------------------------------------ whole-model quantisation
def compute_quantization_error(x_fp32, x_int8, scale) -> Dict:
    """quantize.py:137-156: statistics of ``|x - dequant(q)|`` - max, mean, mean square, and the SNR in dB against the spread
    of that absolute error (per-channel scales broadcast along axis 0)."""
    x = _np(x_fp32).astype(np.float32)
    q = _np(x_int8).astype(np.float32)
    if isinstance(scale, np.ndarray) and scale.ndim > 0:
        deq = q * scale.reshape((len(scale),) + (1,) * (x.ndim - 1))
    else:
        deq = q * scale
    err = np.abs(x - deq)
    return {"max_error": float(np.max(err)), "mean_error": float(np.mean(err)), "mse": float(np.mean(err ** 2)),
            "snr_db": float(20 * np.log10(np.std(x) / (np.std(err) + 1e-12)))}


def quantize_model_per_channel(named_params: Dict[str, np.ndarray]) -> Dict:
    """quantize.py:162-214 on the GPU quantisers: ``named_params`` maps ``"<layer>.weight"`` / ``"<layer>.bias"`` to float
    arrays (what ``module.weight.data`` / ``module.bias.data`` hold; the reference walks an ``nn.Module``, which this path does
    not need).  Weights: per output channel (:func:`quantize_symmetric_per_channel`); biases: per tensor
    (:func:`quantize_symmetric_per_tensor`).  Same result dictionary."""
    out: Dict[str, Dict] = {}
    for key, val in named_params.items():
        arr = _np(val).astype(np.float32)
        if key.endswith(".weight"):
            q, sc = quantize_symmetric_per_channel(arr, axis=0)
            out[key] = {"data": q, "scales": sc, "shape": arr.shape, "axis": 0, "error": compute_quantization_error(arr, q, sc)}
        elif key.endswith(".bias"):
            q, sc = quantize_symmetric_per_tensor(arr)
            out[key] = {"data": q, "scale": sc, "shape": arr.shape, "error": compute_quantization_error(arr, q, sc)}
    return out


# ------------------------------------------------------------------------------------------ 200-byte block variant
PADDED_BLOCK_BYTES = 200      # 25 x 8 bytes: the AXI-burst-aligned alternative discussed at export_bsr_14x14.py:17-21


def save_bsr_binary_int8_padded200(bsr_data: Dict, filepath: str) -> None:
    """The 64-bit aligned variant of the 14x14 block file that export_bsr_14x14.py:17-21 describes and leaves unbuilt: every
    196-byte block followed by 4 zero bytes (200 = 25 x 8).  Same block order as :func:`save_bsr_binary_int8`."""
    data = _np(bsr_data["data"])
    if data.dtype != np.int8:
        raise ValueError(f"Expected INT8 data, got {data.dtype}")
    nnz = int(data.shape[0])
    out = np.zeros((nnz, PADDED_BLOCK_BYTES), dtype=np.int8)
    out[:, :BLOCK_ELEMENTS] = data.reshape(nnz, -1)
    with open(filepath, "wb") as f:
        f.write(out.tobytes())


def load_bsr_binary_int8(filepath: str, padded: bool = False) -> np.ndarray:
    """Reader for ``weights.bsr`` of the 14x14 export (raw 196-byte blocks, export_bsr_14x14.py:241-272) and of its
    200-byte-padded variant.  Returns int8 [nnz, 14, 14]; raises ``ValueError`` when the size is not a whole number of blocks."""
    flat = np.fromfile(filepath, dtype=np.int8)
    stride = PADDED_BLOCK_BYTES if padded else BLOCK_ELEMENTS
    if flat.size % stride:
        raise ValueError(f"{filepath}: {flat.size} bytes is not a multiple of {stride}")
    blocks = flat.reshape(-1, stride)
    if padded and blocks[:, BLOCK_ELEMENTS:].any():
        raise ValueError(f"{filepath}: non-zero padding bytes")
    return np.ascontiguousarray(blocks[:, :BLOCK_ELEMENTS]).reshape(-1, BLOCK_H, BLOCK_W)
