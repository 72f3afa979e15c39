"""Host driver entry points of the reference, dispatching to the B200.

Mirrors ``sw/host/accel.py`` (``AccelDriver``) and ``sw/host/memory.py`` (``BSRMatrix``,
``pack_activations``): same method names, argument meaning, assertions and return shapes, so code
written against the PYNQ driver runs unchanged.  Where the FPGA driver moved bytes over AXI DMA and
polled CSRs, this one keeps weights resident in HBM as MMA tiles and launches one kernel.

Additions over the reference (which never returns the output matrix, only four result registers):
``read_output()``, and optional ``per_channel_scales`` / ``relu`` / ``bias`` / ``residual`` on
``run_inference``.
"""
from __future__ import annotations

import time
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops
from ._lib import AcceleratorError


def channel_scale_factors(scale_act: float, scale_w, scale_out: float) -> np.ndarray:
    """Per-channel requant factor: float32 product, then float32 divide - the
    channel-wise extension of ``in_scale / out_scale`` in requantize_int32_to_int8
    (hw/sim/cpp/src/golden_models.cpp:384)."""
    prod = (np.float32(scale_act) * np.asarray(scale_w, dtype=np.float32)).astype(np.float32)
    return (prod / np.float32(scale_out)).astype(np.float32)


class BSRMatrix:
    """sw/host/memory.py:92-257."""

    def __init__(self, block_size: int = 14):
        self.block_size = block_size
        self.row_ptr = None
        self.col_idx = None
        self.values = None
        self.shape = (0, 0)
        self.nnz_blocks = 0

    @classmethod
    def from_dense(cls, dense: np.ndarray, block_size: int = 14, threshold: float = 0.0) -> "BSRMatrix":
        """memory.py:121-165.  Note the reference's default keeps EVERY block (``norm >= 0.0``)."""
        w = ops.to_device(dense, torch.int8 if np.asarray(dense).dtype == np.int8 else torch.float32)
        norms = ops.block_l2_f32(w.to(torch.float32), block_size, block_size)
        keep = norms >= float(threshold)
        if w.dtype == torch.int8:
            rp, ci, blocks = ops.pack_bsr_i8(w, keep, block_size)
        else:
            rp, ci, blocks = ops.pack_bsr_f32(w, keep, block_size, block_size)
        m = cls(block_size)
        m.row_ptr = rp.cpu().numpy().astype(np.uint32)
        m.col_idx = ci.cpu().numpy().astype(np.uint16)
        m.values = blocks.cpu().numpy().astype(np.asarray(dense).dtype)
        bs = block_size
        m.shape = (-(-dense.shape[0] // bs) * bs, -(-dense.shape[1] // bs) * bs)
        m.nnz_blocks = int(ci.numel())
        return m

    def to_dense(self) -> np.ndarray:
        K, N = self.shape
        bs = self.block_size
        dense = np.zeros((K, N), dtype=self.values.dtype)
        for row in range(len(self.row_ptr) - 1):
            for ptr in range(int(self.row_ptr[row]), int(self.row_ptr[row + 1])):
                col = int(self.col_idx[ptr])
                dense[row * bs:(row + 1) * bs, col * bs:(col + 1) * bs] = self.values[ptr]
        return dense

    def pack_for_dma(self) -> bytes:
        """memory.py:208-222: [row_ptr u32][col_idx u16][values int8]."""
        return self.row_ptr.tobytes() + self.col_idx.tobytes() + self.values.astype(np.int8).tobytes()

    def memory_size(self) -> int:
        return len(self.row_ptr) * 4 + len(self.col_idx) * 2 + self.nnz_blocks * self.block_size * self.block_size

    def sparsity(self) -> float:
        K, N = self.shape
        total = (K // self.block_size) * (N // self.block_size)
        return 0.0 if total == 0 else 1.0 - self.nnz_blocks / total

    def __repr__(self) -> str:
        return (f"BSRMatrix(shape={self.shape}, block_size={self.block_size}, "
                f"nnz_blocks={self.nnz_blocks}, sparsity={self.sparsity():.1%})")


def pack_activations(activations: np.ndarray, block_size: int = 14) -> bytes:
    """memory.py:260-282: zero-pad [M, K] int8 to multiples of the block, row-major bytes."""
    M, K = activations.shape
    Mp, Kp = -(-M // block_size) * block_size, -(-K // block_size) * block_size
    if (Mp, Kp) != (M, K):
        padded = np.zeros((Mp, Kp), dtype=np.int8)
        padded[:M, :K] = activations
        activations = padded
    return activations.tobytes()


class AccelDriver:
    """Drop-in for ``sw/host/accel.py::AccelDriver`` (lines 102-435) backed by libaccel_b200.so.

    Dimension naming follows accel.py: M = activation rows, N = output columns, K = reduction.
    Weights are Convention B (block-rows = output channels), as every reference exporter writes them.
    Not thread-safe, like the reference (accelerator_driver.hpp:44-47).
    """

    BLOCK_SIZE = 14
    DATA_WIDTH = 8
    ACC_WIDTH = 32

    def __init__(self, overlay=None, csr_base: int = 0x43C00000, dma_base: int = 0x40000000,
                 simulation: bool = False):
        # overlay / csr_base / dma_base / simulation are accepted for signature compatibility; there is
        # no register file or simulated CSR here - the device is the GPU or nothing.
        self.csr_base, self.dma_base, self.simulation = csr_base, dma_base, False
        self.M = self.N = self.K = 0
        self.weights_loaded = False
        self.activations_loaded = False
        self._plan: Optional[ops.BsrPlan] = None
        self._act: Optional[torch.Tensor] = None
        self._out: Optional[torch.Tensor] = None
        self._scales: Optional[Tuple[float, float]] = None
        self._use_dense = False
        self._perf = {"total_cycles": 0, "active_cycles": 0, "idle_cycles": 0, "cache_hits": 0, "cache_misses": 0}

    def configure_dimensions(self, M: int, N: int, K: int, Tm: int = 14, Tn: int = 14, Tk: int = 14):
        self.M, self.N, self.K = int(M), int(N), int(K)
        self.Tm, self.Tn, self.Tk = Tm, Tn, Tk

    def load_sparse_weights(self, row_ptr, col_idx, weights, block_size: int = 14) -> int:
        """accel.py:177-236.  Returns the byte count of the reference's DMA blob (u32|u16|int8)."""
        assert block_size == self.BLOCK_SIZE, f"Block size must be {self.BLOCK_SIZE}"
        w_dtype = weights.dtype
        assert w_dtype in (np.int8, torch.int8), "Weights must be INT8"
        nbc = -(-self.K // block_size) if self.K else None
        if nbc is not None and len(col_idx):
            nbc = max(nbc, int(np.max(np.asarray(col_idx))) + 1)
        self._plan = ops.BsrPlan(np.asarray(row_ptr).astype(np.int64), np.asarray(col_idx).astype(np.int64), weights,
                                 n_block_cols=nbc)
        self.weights_loaded = True
        return 4 * len(row_ptr) + 2 * len(col_idx) + int(np.prod(weights.shape))

    def load_activations(self, activations) -> int:
        """accel.py:238-277: int8 [M, K]; host arrays are copied to the device, CUDA tensors are used in place."""
        assert activations.dtype in (np.int8, torch.int8), "Activations must be INT8"
        assert tuple(activations.shape) == (self.M, self.K), \
            f"Activation shape {tuple(activations.shape)} doesn't match (M={self.M}, K={self.K})"
        self._act = ops.to_device(activations, torch.int8)
        self.activations_loaded = True
        return self.M * self.K

    def set_scale_factors(self, Sa: float, Sw: float):
        self._scales = (float(Sa), float(Sw))

    def set_scheduler_mode(self, use_dense: bool):
        self._use_dense = bool(use_dense)      # both schedulers compute the same result; kept for API parity

    def run_inference(self, timeout_ms: int = 1000, per_channel_scales=None, relu: bool = False, bias=None,
                      residual=None, res_scales=None) -> Tuple[bool, dict]:
        """accel.py:279-342.  Launches the BSR GEMM; INT32 output unless ``per_channel_scales`` is given."""
        assert self.weights_loaded, "Weights not loaded"
        assert self.activations_loaded, "Activations not loaded"
        n_pad = self._plan.n_out_padded
        n = min(self.N, n_pad) if self.N else n_pad
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = time.time()
        try:
            ev0.record()
            if per_channel_scales is None:
                self._out = self._plan.gemm(self._act, "i32", n_channels=n, bias=bias, relu=relu)
            else:
                self._out = self._plan.gemm(self._act, "i8", n_channels=n, chan_scale=per_channel_scales, bias=bias,
                                            relu=relu, residual=residual, res_scales=res_scales)
            ev1.record()
            while not ev1.query():
                if (time.time() - t_host) * 1000.0 > timeout_ms:
                    return False, {"error": "timeout"}
                time.sleep(0.0001)
        except AcceleratorError as e:
            return False, {"error": str(e)}
        ms = ev0.elapsed_time(ev1)
        clk_khz = torch.cuda.get_device_properties(self._act.device).clock_rate if hasattr(
            torch.cuda.get_device_properties(self._act.device), "clock_rate") else 1_965_000
        cycles = int(ms * clk_khz)
        self._perf.update(total_cycles=cycles, active_cycles=cycles, idle_cycles=0)
        sample = self._out.reshape(-1)[:4].to(torch.int64).cpu().tolist()
        sample += [0] * (4 - len(sample))
        return True, {"cycles": cycles, "active_cycles": cycles, "utilization": 100.0 if cycles else 0.0,
                      "result_sample": sample, "error": 0, "elapsed_ms": ms}

    def read_output(self):
        """Full output matrix of the last run (the reference exposes only RESULT_0..3)."""
        assert self._out is not None, "run_inference has not been called"
        return self._out.cpu().numpy()

    def get_performance_stats(self) -> dict:
        return dict(self._perf)

    def reset(self):
        self.weights_loaded = False
        self.activations_loaded = False
        self._plan = self._act = self._out = None
