"""MNIST-CNN INT8 pipeline with a BSR-pruned FC1 (BASELINE config 1) on the B200 path.

Topology (identical in ``sw/training/export_bsr_14x14.py:60-78``, ``sw/INT8 quantization/quantize.py:34-49``,
``sw/training/blocksparse_train.py:31-49``)::

    conv1 1->32 3x3 -> ReLU -> conv2 32->64 3x3 -> ReLU -> maxpool 2x2 -> flatten (C,H,W) -> fc1 9216->128 -> ReLU -> fc2 128->10

Inputs are the reference's own artefacts: ``<layer>_weight_int8.npy`` / ``_weight_scales.npy`` / ``_bias_int8.npy`` /
``_bias_scale.json`` (``data/int8``, written by ``quantize.py:162-208``).  What the reference leaves open is defined in
SURVEY.md A.7 and implemented here (the oracle holds the CPU twin):

* input quantisation: per-tensor symmetric, ``q = clip(rint(x / s_x))`` with ``s_x`` from the calibration batch;
* INT32 bias in the accumulator domain: ``rint((bias_int8 * bias_scale) / (s_act * s_w[c]))`` (float64, host);
* per-channel requant ``sf[c] = s_act * s_w[c] / s_out`` fused into every layer's epilogue;
* logits: INT32 accumulators of fc2 de-quantised per channel, ``float(acc) * s_fc1_out * s_w[c]``, for the arg-max.

FC1 is pruned to the requested block sparsity on the GPU: block L2 norms of the INT8 weights, the selection rule of
``prune_blocks_global`` (``blocksparse_train.py:141-239``, one layer, no floor), packed by the GPU packer.  Every layer runs
through the C ABI (``accel_conv_bsr_i8``, ``accel_maxpool_i8``, ``accel_bsr_gemm_i8``); the whole forward is captured into one
CUDA graph.  No CPU fallback.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import exporters, ops
from .host import channel_scale_factors

MNIST_MEAN, MNIST_STD = 0.1307, 0.3081
LAYERS = ("conv1", "conv2", "fc1", "fc2")


def load_int8_dir(int8_dir: str) -> Dict[str, np.ndarray]:
    """Read the reference's ``data/int8`` layout (quantize.py:185-208)."""
    w: Dict[str, np.ndarray] = {}
    for n in LAYERS:
        w[f"{n}_weight_int8"] = np.load(os.path.join(int8_dir, f"{n}_weight_int8.npy"))
        w[f"{n}_weight_scales"] = np.load(os.path.join(int8_dir, f"{n}_weight_scales.npy")).astype(np.float32)
        w[f"{n}_bias_int8"] = np.load(os.path.join(int8_dir, f"{n}_bias_int8.npy"))
        with open(os.path.join(int8_dir, f"{n}_bias_scale.json")) as f:
            js = json.load(f)
        w[f"{n}_bias_scale"] = np.float64(js["scale"] if isinstance(js, dict) else js)
    return w


def preprocess(images_u8, mode: str = "normalized") -> torch.Tensor:
    """uint8 [B, 28, 28] -> float32 CUDA [B, 1, 28, 28].  "normalized": quantize.py:226-235; "raw": the pixel values, which is
    how the shipped golden logits were produced (train_mnist.py:161-166)."""
    x = ops.to_device(images_u8, torch.uint8).to(torch.float32)
    if mode == "normalized":
        # IEEE float32 divisions, as NumPy does them: the divisors are device tensors on purpose (with a host scalar PyTorch
        # multiplies by the reciprocal, which rounds differently)
        dev = x.device
        x = x / torch.tensor(255.0, dtype=torch.float32, device=dev)
        x = (x - torch.tensor(MNIST_MEAN, dtype=torch.float32, device=dev)) / torch.tensor(MNIST_STD, dtype=torch.float32, device=dev)
    elif mode != "raw":
        raise ValueError(mode)
    return x[:, None, :, :].contiguous()


class ActivationCalibrator:
    """Max-abs calibration of the per-tensor activation scales (``quantize_activations_from_golden``, quantize.py:217-264;
    the forward-hook min/max collector of ``quantize_resnet18.py:103-160`` has the same contract).  ``observe`` may be
    called once per calibration batch; ``scales()`` returns ``max(maxabs / 127, 1e-12)`` per tensor."""

    def __init__(self):
        self.maxabs: Dict[str, float] = {}

    def observe(self, name: str, t: torch.Tensor) -> None:
        m = float(t.detach().abs().max().item()) if t.numel() else 0.0
        self.maxabs[name] = max(self.maxabs.get(name, 0.0), m)

    def scales(self) -> Dict[str, float]:
        return {k: max(v / 127.0, 1e-12) for k, v in self.maxabs.items()}


class MnistCnnInt8:
    """The quantised MNIST CNN on the GPU.  ``build`` prunes / packs / uploads, ``calibrate`` derives the activation scales
    from a float forward of the de-quantised weights, ``run`` returns de-quantised logits [B, 10] (CUDA float32)."""

    def __init__(self, weights: Dict[str, np.ndarray], batch: int = 64, fc1_sparsity: float = 0.9, mode: str = "normalized"):
        ops._require_cuda()
        self.w, self.batch, self.fc1_sparsity, self.mode = weights, int(batch), float(fc1_sparsity), mode
        self.scales: Optional[Dict[str, float]] = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.plans: Dict[str, ops.BsrPlan] = {}
        self.bsr: Dict[str, Dict] = {}
        self.sat = torch.zeros(1, dtype=torch.int64, device="cuda")

    # ------------------------------------------------------------------ calibration (float forward, plumbing: torch ops)
    def float_forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        import torch.nn.functional as F

        def deq(n):
            wq = torch.from_numpy(self.w[f"{n}_weight_int8"]).cuda().to(torch.float32)
            s = torch.from_numpy(self.w[f"{n}_weight_scales"]).cuda().reshape((-1,) + (1,) * (wq.dim() - 1))
            b = torch.from_numpy(self.w[f"{n}_bias_int8"].astype(np.float32)).cuda() * float(self.w[f"{n}_bias_scale"])
            return wq * s, b
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            (w1, b1), (w2, b2), (w3, b3), (w4, b4) = (deq(n) for n in LAYERS)
            a1 = F.relu(F.conv2d(x, w1, b1))
            a2 = F.relu(F.conv2d(a1, w2, b2))
            f1 = F.relu(F.linear(F.max_pool2d(a2, 2).flatten(1), w3, b3))
            logits = F.linear(f1, w4, b4)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        return {"input": x, "conv1_out": a1, "conv2_out": a2, "fc1_out": f1, "logits": logits}

    def calibrate(self, images_u8) -> Dict[str, float]:
        acts = self.float_forward(preprocess(images_u8, self.mode))
        cal = ActivationCalibrator()
        for k in ("input", "conv1_out", "conv2_out", "fc1_out"):
            cal.observe(k, acts[k])
        self.scales = cal.scales()
        return self.scales

    # ------------------------------------------------------------------ build
    def prune_fc1(self) -> Dict:
        """Block-L2 pruning of FC1 on the GPU (one layer, no floor), then the GPU packer."""
        w = torch.from_numpy(self.w["fc1_weight_int8"]).cuda()
        if self.fc1_sparsity <= 0.0:
            return exporters.build_bsr_14x14_int8_direct(w, device=True)
        norms, _, _ = exporters.compute_block_norms(w.to(torch.float32), exporters.BLOCK_H, exporters.BLOCK_W)
        keep = exporters.prune_blocks_global([norms], self.fc1_sparsity, [0.0])[0]
        full = keep.repeat_interleave(exporters.BLOCK_H, 0).repeat_interleave(exporters.BLOCK_W, 1)[: w.shape[0], : w.shape[1]]
        return exporters.build_bsr_14x14_int8_direct(w * full.to(torch.int8), device=True)

    def build(self, scales: Optional[Dict[str, float]] = None) -> None:
        if scales is not None:
            self.scales = dict(scales)
        if self.scales is None:
            raise RuntimeError("activation scales missing: call calibrate(images) or pass scales")
        s_x, s1, s2, s3 = (self.scales[k] for k in ("input", "conv1_out", "conv2_out", "fc1_out"))
        for n in LAYERS:
            if n == "fc1":
                self.bsr[n] = self.prune_fc1()
            else:
                wq = self.w[f"{n}_weight_int8"]
                self.bsr[n] = exporters.build_bsr_14x14_int8_direct(torch.from_numpy(wq.reshape(wq.shape[0], -1)).cuda(), device=True)
            b = self.bsr[n]
            self.plans[n] = ops.BsrPlan(b["indptr"], b["indices"], b["data"], n_block_cols=b["num_block_cols"])
        sw = {n: self.w[f"{n}_weight_scales"].astype(np.float32) for n in LAYERS}
        self.sf = {"conv1": channel_scale_factors(s_x, sw["conv1"], s1), "conv2": channel_scale_factors(s1, sw["conv2"], s2),
                   "fc1": channel_scale_factors(s2, sw["fc1"], s3)}
        self.sf = {k: torch.from_numpy(v).cuda() for k, v in self.sf.items()}
        self.logit_scale = torch.from_numpy((np.float32(s3) * sw["fc2"]).astype(np.float32)).cuda()
        self.bias = {}
        for n, s_in in (("conv1", s_x), ("conv2", s1), ("fc1", s2), ("fc2", s3)):
            real = self.w[f"{n}_bias_int8"].astype(np.float64) * float(self.w[f"{n}_bias_scale"])
            self.bias[n] = torch.from_numpy(np.rint(real / (float(s_in) * sw[n].astype(np.float64))).astype(np.int32)).cuda()
        B = self.batch
        self.x_q = torch.zeros((B, 1, 28, 28), dtype=torch.int8, device="cuda")
        self.buf = {"conv1": torch.empty((B, 32, 26, 26), dtype=torch.int8, device="cuda"),
                    "conv2": torch.empty((B, 64, 24, 24), dtype=torch.int8, device="cuda"),
                    "pool": torch.empty((B, 64, 12, 12), dtype=torch.int8, device="cuda"),
                    "fc1": torch.empty((B, 128), dtype=torch.int8, device="cuda"),
                    "logits_i32": torch.empty((B, 10), dtype=torch.int32, device="cuda"),
                    "logits": torch.empty((B, 10), dtype=torch.float32, device="cuda")}
        self.graph = None

    # ------------------------------------------------------------------ forward
    def quantize_input(self, images_u8) -> torch.Tensor:
        """clip(rint(x / s_x), -128, 127) on the GPU (the per-row quantiser kernel with one row)."""
        x = preprocess(images_u8, self.mode)
        s = torch.full((1,), np.float32(self.scales["input"]), dtype=torch.float32, device="cuda")
        return ops.quantize_rows_f32(x.reshape(1, -1), s).reshape(x.shape)

    def forward(self) -> torch.Tensor:
        """Enqueue the five launches on the current stream (reads ``self.x_q``)."""
        P, b = self.plans, self.buf
        P["conv1"].conv(self.x_q, 3, 1, 0, 32, "i8", chan_scale=self.sf["conv1"], bias=self.bias["conv1"], relu=True,
                        out=b["conv1"], sat_count=self.sat)
        P["conv2"].conv(b["conv1"], 3, 1, 0, 64, "i8", chan_scale=self.sf["conv2"], bias=self.bias["conv2"], relu=True,
                        out=b["conv2"], sat_count=self.sat)
        ops.maxpool_i8(b["conv2"], 2, 2, 0, out=b["pool"])
        P["fc1"].gemm(b["pool"].reshape(self.batch, -1), "i8", n_channels=128, chan_scale=self.sf["fc1"], bias=self.bias["fc1"],
                      relu=True, out=b["fc1"], sat_count=self.sat)
        P["fc2"].gemm(b["fc1"], "i32", n_channels=10, bias=self.bias["fc2"], out=b["logits_i32"])
        P["fc2"].gemm(b["fc1"], "f32", n_channels=10, chan_scale=self.logit_scale, bias=self.bias["fc2"], out=b["logits"])
        return b["logits"]

    def capture(self) -> None:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward()
        self.graph = g

    def run(self, images_u8) -> torch.Tensor:
        """uint8 [batch, 28, 28] (host or device) -> de-quantised logits float32 [batch, 10] (CUDA)."""
        if not self.plans:
            raise RuntimeError("Weights not loaded")
        xq = self.quantize_input(images_u8)
        if tuple(xq.shape) != tuple(self.x_q.shape):
            raise ValueError(f"expected {self.batch} images of 28x28, got {tuple(xq.shape)}")
        self.x_q.copy_(xq)
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.buf["logits"]

    # ------------------------------------------------------------------ accounting
    def work(self) -> Dict:
        B = self.batch
        rows = {"conv1": B * 26 * 26, "conv2": B * 24 * 24, "fc1": B, "fc2": B}
        in_b = {"conv1": B * 28 * 28, "conv2": B * 32 * 26 * 26, "fc1": B * 9216, "fc2": B * 128}
        out_b = {"conv1": B * 32 * 26 * 26, "conv2": B * 64 * 24 * 24, "fc1": B * 128, "fc2": B * 10 * 4}
        o = sum(2 * rows[n] * self.plans[n].num_blocks * 196 for n in LAYERS)
        by = sum(in_b[n] + out_b[n] + self.plans[n].num_blocks * 200 + 4 * (self.plans[n].n_block_rows + 1) for n in LAYERS)
        by += B * 64 * 24 * 24 + B * 64 * 12 * 12          # max-pool read + write
        return {"ops": o, "bytes": by}
