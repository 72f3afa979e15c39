"""Layer objects, layer tables and the network runner for the BSR-INT8 path.

* ``ConvSpec`` / ``resnet18_specs`` - the reference's ResNet-18 layer table
  (hw/sim/cpp/src/resnet_inference.cpp:61-127) with the three 1x1 downsample convolutions the
  reference only flags (``export_resnet18_bsr.py:72,79,86``) written out as layers of their own,
  plus the 3x3/2 stem max-pool the table omits.
* ``synthetic_conv_weights`` - the reference's synthetic-weight recipe (He init of
  ``create_conv_weights`` sw/exporters/export_conv.py:17-40, block mask of ``create_sparse_mask``
  with one seed per layer as ``BlockSparsePruner._create_masks`` train_resnet18.py:163-184,
  per-channel INT8 quantisation quantize.py:71-98) with 14x14 blocks throughout.
* ``BsrConv`` / ``BsrLinear`` / ``BsrNetwork`` - device-resident layers that call the C ABI, and a
  runner that replays the whole network as one CUDA graph (``ResNetInference::run_inference``
  resnet_inference.hpp:180-271 is the reference-side shape of that API).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import exporters, ops
from .host import channel_scale_factors


@dataclass
class ConvSpec:
    name: str
    c_in: int
    c_out: int
    h: int            # input height
    w: int            # input width
    k: int
    stride: int
    pad: int
    relu: bool = True           # ReLU on the INT32 accumulator (no residual) / on the int8 sum (residual)
    residual: Optional[str] = None   # name of the tensor added after requant ("input of the block" or a downsample)
    src: Optional[str] = None        # input tensor name (default: previous layer's output)
    kind: str = "conv"               # conv | maxpool | avgpool | fc

    @property
    def h_out(self) -> int:
        return (self.h + 2 * self.pad - self.k) // self.stride + 1

    @property
    def w_out(self) -> int:
        return (self.w + 2 * self.pad - self.k) // self.stride + 1

    @property
    def K(self) -> int:
        return self.c_in * self.k * self.k


def resnet18_specs(image: int = 224, num_classes: int = 1000) -> List[ConvSpec]:
    s: List[ConvSpec] = []
    h = image
    s.append(ConvSpec("conv1", 3, 64, h, h, 7, 2, 3))
    h = s[-1].h_out
    s.append(ConvSpec("maxpool", 64, 64, h, h, 3, 2, 1, kind="maxpool"))
    h = s[-1].h_out
    c = 64
    for stage, width in enumerate((64, 128, 256, 512), start=1):
        for blk in range(2):
            stride = 2 if (stage > 1 and blk == 0) else 1
            block_in = s[-1].name
            pre = f"layer{stage}.{blk}"
            s.append(ConvSpec(f"{pre}.conv1", c, width, h, h, 3, stride, 1, src=block_in))
            h_out = s[-1].h_out
            ident = block_in
            if stride != 1 or c != width:
                s.append(ConvSpec(f"{pre}.downsample", c, width, h, h, 1, stride, 0, relu=False, src=block_in))
                ident = s[-1].name
            s.append(ConvSpec(f"{pre}.conv2", width, width, h_out, h_out, 3, 1, 1, residual=ident,
                              src=f"{pre}.conv1"))
            h, c = h_out, width
    s.append(ConvSpec("avgpool", c, c, h, h, h, 1, 0, kind="avgpool"))
    s.append(ConvSpec("fc", c, num_classes, 1, 1, 1, 1, 0, relu=False, kind="fc"))
    return s


def resnet50_specs(image: int = 224, num_classes: int = 1000) -> List[ConvSpec]:
    """torchvision resnet50 (v1.5: stride on the 3x3), SURVEY.md A.9."""
    s: List[ConvSpec] = []
    h = image
    s.append(ConvSpec("conv1", 3, 64, h, h, 7, 2, 3))
    h = s[-1].h_out
    s.append(ConvSpec("maxpool", 64, 64, h, h, 3, 2, 1, kind="maxpool"))
    h = s[-1].h_out
    c = 64
    for stage, (width, n) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for blk in range(n):
            stride = 2 if (stage > 1 and blk == 0) else 1
            block_in = s[-1].name
            pre = f"layer{stage}.{blk}"
            s.append(ConvSpec(f"{pre}.conv1", c, width, h, h, 1, 1, 0, src=block_in))
            s.append(ConvSpec(f"{pre}.conv2", width, width, h, h, 3, stride, 1))
            h_out = s[-1].h_out
            ident = block_in
            if stride != 1 or c != 4 * width:
                s.append(ConvSpec(f"{pre}.downsample", c, 4 * width, h, h, 1, stride, 0, relu=False, src=block_in))
                ident = s[-1].name
            s.append(ConvSpec(f"{pre}.conv3", width, 4 * width, h_out, h_out, 1, 1, 0, residual=ident,
                              src=f"{pre}.conv2"))
            h, c = h_out, 4 * width
    s.append(ConvSpec("avgpool", c, c, h, h, h, 1, 0, kind="avgpool"))
    s.append(ConvSpec("fc", c, num_classes, 1, 1, 1, 1, 0, relu=False, kind="fc"))
    return s


# Synthetic quantisation constants of SURVEY.md 8d: fixed activation scales so that requant saturates sometimes.
S_ACT_IN = 0.02
S_ACT_OUT = 0.05


def synthetic_conv_weights(spec: ConvSpec, sparsity_pct: float, layer_idx: int, bias_range: int = 0) -> Dict:
    """FP32 He weights -> 14x14 block mask -> per-channel INT8 (host RNG streams as the reference)."""
    seed = 42 + layer_idx
    np.random.seed(seed)
    scale = np.sqrt(2.0 / (spec.c_in * spec.k * spec.k))
    w4 = np.random.randn(spec.c_out, spec.c_in, spec.k, spec.k).astype(np.float32) * scale
    w2 = w4.reshape(spec.c_out, -1)
    mask = exporters.create_sparse_mask(w2.shape, sparsity_pct, block_size=exporters.BLOCK_SIZE, seed=seed)
    w2 = (w2 * mask).astype(np.float32)
    bias = None
    if bias_range:
        bias = np.random.default_rng(seed).integers(-bias_range, bias_range + 1, spec.c_out).astype(np.int32)
    return {"w2": w2, "bias": bias}


class BsrLayer:
    """One BSR weight matrix on the device + its epilogue constants."""

    def __init__(self, spec: ConvSpec, w_fp32_2d=None, *, bsr: Optional[Dict] = None, w_scales=None, bias=None,
                 s_in: float = S_ACT_IN, s_out: float = S_ACT_OUT, group_rows: int = 0):
        self.spec = spec
        if bsr is None:
            q, w_scales = exporters.quantize_symmetric_per_channel(w_fp32_2d, device=True)
            bsr = exporters.build_bsr_14x14_int8_direct(q, device=True)
        self.bsr = bsr
        self.plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"],
                                group_rows=group_rows)
        ws = w_scales.cpu().numpy() if isinstance(w_scales, torch.Tensor) else np.asarray(w_scales, np.float32)
        self.w_scales = ws.astype(np.float32)
        self.s_in, self.s_out = float(s_in), float(s_out)
        sf_host = channel_scale_factors(s_in, self.w_scales, s_out)
        self.sf = torch.from_numpy(sf_host).cuda()
        # requant is monotone in the accumulator only for positive factors: decided here, on the host, once - the fused
        # conv + max-pool kernel pools on INT32 accumulators and must not be taken otherwise (no device sync on the forward path)
        self.sf_positive = bool((sf_host[: spec.c_out] > 0).all())
        self.bias = None if bias is None else torch.as_tensor(bias, dtype=torch.int32).cuda()

    # algorithmic work of one call
    def useful_ops(self, M: int) -> int:
        return 2 * M * self.plan.num_blocks * 196

    def weight_bytes(self) -> int:
        return self.plan.num_blocks * 200 + 4 * (self.plan.n_block_rows + 1) + 4 * self.spec.c_out


def input_scales(specs: List[ConvSpec], s_input: float, s_out: float) -> Dict[str, float]:
    """Activation scale of the tensor every conv / fc layer READS when scales chain through the network: the network input
    is quantised with ``s_input``, every convolution requantises to ``s_out``, pools keep the scale of their input."""
    scale_of: Dict[str, float] = {"input": float(s_input)}
    s_in: Dict[str, float] = {}
    prev = "input"
    for sp in specs:
        src = sp.src or prev
        if sp.kind in ("conv", "fc"):
            s_in[sp.name] = scale_of[src]
            scale_of[sp.name] = float(s_out)
        else:
            scale_of[sp.name] = scale_of[src]
        prev = sp.name
    return s_in


class BsrNetwork:
    """A chain of BSR conv / pool / fc layers on one GPU.  ``forward`` enqueues every layer on the current
    stream; ``capture`` records it once into a CUDA graph for replay (launch-latency-bound tails).

    Activation scales: every tensor has ONE scale (``scale_of``).  A layer's requant factor is built from the scale of the
    tensor it reads and its own output scale, the residual add uses (own output scale, scale of the identity tensor, own
    output scale).  ``chain_scales=False`` is the synthetic recipe of SURVEY.md 8d, where every layer is DEFINED to read
    activations at ``S_ACT_IN`` and to write at ``S_ACT_OUT`` regardless of its producer (a throughput / parity workload,
    not a calibrated model); ``chain_scales=True`` makes each layer read at its producer's output scale, which is what a
    real quantised model needs (``ResNetInference.load_model``)."""

    def __init__(self, specs: List[ConvSpec], sparsity_pct: float, batch: int, bias_range: int = 0,
                 layers: Optional[Dict[str, BsrLayer]] = None, s_input: float = S_ACT_IN, s_out: float = S_ACT_OUT,
                 chain_scales: bool = False, shard: Optional[Tuple[int, int]] = None):
        self.specs, self.batch, self.sparsity_pct = specs, batch, sparsity_pct
        self.layers: Dict[str, BsrLayer] = layers or {}
        if not self.layers:
            chained = input_scales(specs, s_input, s_out)
            idx = 0
            for sp in specs:
                if sp.kind in ("conv", "fc"):
                    syn = synthetic_conv_weights(sp, sparsity_pct, idx, bias_range)
                    self.layers[sp.name] = BsrLayer(sp, syn["w2"], bias=syn["bias"],
                                                    s_in=(chained[sp.name] if chain_scales else s_input), s_out=s_out,
                                                    group_rows=(int(__import__('os').environ.get('ACCEL_FC_GROUP_ROWS', 8)) if sp.kind == "fc" else 0))
                    idx += 1
        # one scale per tensor: conv outputs carry their layer's s_out, pools keep their input's scale
        self.scale_of: Dict[str, float] = {"input": float(s_input)}
        prev = "input"
        for sp in specs:
            src = sp.src or prev
            self.scale_of[sp.name] = self.layers[sp.name].s_out if sp.kind in ("conv", "fc") else self.scale_of[src]
            prev = sp.name
        self.buffers: Dict[str, torch.Tensor] = {}
        self.sat = torch.zeros(1, dtype=torch.int64, device="cuda")
        for sp in specs:
            if sp.kind == "fc":
                self.buffers[sp.name] = torch.empty((batch, sp.c_out), dtype=torch.int32, device="cuda")
            elif sp.kind == "avgpool":
                self.buffers[sp.name] = torch.empty((batch, sp.c_out), dtype=torch.int8, device="cuda")
            else:   # rows padded to 16 bytes: the next layer streams them with 16-byte cp.async
                self.buffers[sp.name] = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.sub_buffers: Dict[str, torch.Tensor] = {}     # inputs of 1x1 / stride-2 convolutions after 2x sub-sampling
        self.static_in: Optional[torch.Tensor] = None
        # a 3x3 / stride 2 convolution and the 1x1 / stride 2 downsample of the same tensor run as one call
        self.fused_ds: Dict[str, str] = {}
        by_name = {sp.name: sp for sp in specs}
        prev = "input"
        for sp in specs:
            if sp.kind == "conv" and sp.k == 1 and sp.stride == 2 and sp.pad == 0 and not sp.residual:
                for cand in specs:
                    if (cand.kind == "conv" and cand.k == 3 and cand.stride == 2 and cand.pad == 1 and not cand.residual
                            and (cand.src or None) == (sp.src or None) and cand.src and cand.c_in == sp.c_in
                            and cand.c_out == sp.c_out and cand.name not in self.fused_ds
                            and specs.index(cand) < specs.index(sp)):
                        self.fused_ds[cand.name] = sp.name
                        break
            prev = sp.name
        # the stem convolution and the max-pool behind it run as one kernel when the library has one for the geometry
        # (csrc/stem_ws.cuh); the convolution's own output tensor is then never materialised
        self.fused_pool: Dict[str, str] = {}
        for i, sp in enumerate(specs[:-1]):
            nx = specs[i + 1]
            if (sp.kind == "conv" and (sp.k, sp.stride, sp.pad) == (7, 2, 3) and sp.relu and not sp.residual
                    and nx.kind == "maxpool" and (nx.k, nx.stride, nx.pad) == (3, 2, 1) and not nx.src
                    and sp.c_in * 7 <= 32 and sp.c_out <= 64 and sp.w % 32 == 0 and sp.w <= 224 and sp.h % 4 == 0
                    and not any(o.src == sp.name or o.residual == sp.name for o in specs)
                    and self.layers[sp.name].sf_positive):
                self.fused_pool[sp.name] = nx.name
                self.buffers.pop(sp.name, None)
        self.n_launches = sum(1 for _ in specs) - len(self.fused_ds) - len(self.fused_pool)

    def forward(self, x: torch.Tensor, launch_events: Optional[list] = None) -> torch.Tensor:
        """Enqueue every layer on the current stream.  ``launch_events``: when a list is passed, one
        ``(layer name, start event, end event)`` triple per kernel launch is appended (per-kernel timing for the roofline
        table; the events are recorded on the launching stream)."""
        prev = "input"
        t: Dict[str, torch.Tensor] = {"input": x}
        t.update(self.buffers)
        pending = None
        for sp in self.specs:
            if launch_events is not None:
                if pending is not None:
                    pending[2].record()
                    launch_events.append(tuple(pending))
                    pending = None
                launched = not (sp.name in self.fused_pool.values() or (sp.kind == "conv" and sp.name in self.fused_ds.values()))
                if launched:
                    pending = [sp.name, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
                    pending[1].record()
            if sp.name in self.fused_pool.values():
                prev = sp.name                          # written by the stem convolution it is fused with
                continue
            src = t[sp.src] if sp.src else t[prev]
            if sp.name in self.fused_pool:
                L = self.layers[sp.name]
                ops.conv_pool(L.plan, src, sp.c_out, chan_scale=L.sf, bias=L.bias, relu=True,
                              out=self.buffers[self.fused_pool[sp.name]], sat_count=self.sat, fused_ok=L.sf_positive)
                prev = sp.name
                continue
            out = self.buffers[sp.name]
            if sp.kind == "conv" and sp.name in self.fused_ds.values():
                pass                                   # written by the stride-2 convolution it is fused with
            elif sp.kind == "conv" and sp.name in self.fused_ds:
                L, D = self.layers[sp.name], self.layers[self.fused_ds[sp.name]]
                dsp = D.spec
                ops.conv_dual(L.plan, D.plan, src, sp.c_out, chan_scale=L.sf, chan_scale_ds=D.sf, bias=L.bias, bias_ds=D.bias,
                              relu=sp.relu, relu_ds=dsp.relu, out=out, out_ds=self.buffers[dsp.name], sat_count=self.sat)
            elif sp.kind == "conv":
                L = self.layers[sp.name]
                stride = sp.stride
                if sp.k == 1 and sp.stride == 2 and sp.pad == 0:
                    # a 1x1 / stride 2 convolution (ResNet-50 downsample) reads every second pixel of every second row: take
                    # those first, then it is a stride-1 pointwise convolution and runs on the weight-stationary kernel
                    if sp.name not in self.sub_buffers:
                        self.sub_buffers[sp.name] = ops.alloc_padded((src.shape[0], sp.c_in, sp.h_out, sp.w_out))
                    src = ops.subsample2_int8(src, out=self.sub_buffers[sp.name])
                    stride = 1
                if sp.residual:
                    # conv -> requant -> + identity -> ReLU on the int8 sum
                    L.plan.conv(src, sp.k, stride, sp.pad, sp.c_out, "i8", chan_scale=L.sf, bias=L.bias, relu=False,
                                residual=t[sp.residual], res_scales=(L.s_out, self.scale_of[sp.residual], L.s_out), out=out,
                                sat_count=self.sat, relu_out=True)
                else:
                    L.plan.conv(src, sp.k, stride, sp.pad, sp.c_out, "i8", chan_scale=L.sf, bias=L.bias,
                                relu=sp.relu, out=out, sat_count=self.sat)
            elif sp.kind == "maxpool":
                ops.maxpool_i8(src, sp.k, sp.stride, sp.pad, out=out)
            elif sp.kind == "avgpool":
                ops.avgpool_i8(src, out=out)
            elif sp.kind == "fc":
                L = self.layers[sp.name]
                L.plan.gemm(src.reshape(src.shape[0], -1), "i32", n_channels=sp.c_out, bias=L.bias, out=out)
            prev = sp.name
        if launch_events is not None and pending is not None:
            pending[2].record()
            launch_events.append(tuple(pending))
        return self.buffers[self.specs[-1].name]

    def capture(self, x: torch.Tensor) -> None:
        self.static_in = ops.alloc_padded(tuple(x.shape))
        self.static_in.copy_(x)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.forward(self.static_in)          # warm-up (also sets the kernels' smem attributes)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward(self.static_in)
        self.graph = g

    def capture_alt(self) -> None:
        """A second graph of the same network whose first kernel reads ``static_in_alt`` instead of ``static_in`` (every other
        buffer is shared): a host loop that alternates the two graphs can copy batch i+1 straight into the idle input while
        batch i computes, with no device-to-device staging copy (``ResNetInference.run_inference_pipelined``)."""
        if self.graph is None:
            raise RuntimeError("capture() first")
        if getattr(self, "graph_alt", None) is not None:
            return
        self.static_in_alt = ops.alloc_padded(tuple(self.static_in.shape))
        self.static_in_alt.copy_(self.static_in)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward(self.static_in_alt)
        self.graph_alt = g

    def replay(self, x: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None:
            self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.buffers[self.specs[-1].name]

    # ------------------------------------------------------------------ accounting
    def work(self) -> Dict:
        ops_total, bytes_total, per_layer = 0, 0, []
        B = self.batch
        for sp in self.specs:
            in_bytes = B * sp.c_in * sp.h * sp.w
            if sp.kind in ("conv", "fc"):
                L = self.layers[sp.name]
                M = B * sp.h_out * sp.w_out
                o = L.useful_ops(M)
                out_bytes = M * sp.c_out * (4 if sp.kind == "fc" else 1)
                b = in_bytes + L.weight_bytes() + out_bytes + (M * sp.c_out if sp.residual else 0)
            else:
                o = 0
                b = in_bytes + B * sp.c_out * (sp.h_out * sp.w_out if sp.kind == "maxpool" else 1)
            ops_total += o
            bytes_total += b
            dense = 2 * B * sp.h_out * sp.w_out * sp.c_out * sp.c_in * sp.k * sp.k if sp.kind in ("conv", "fc") else 0
            per_layer.append({"name": sp.name, "ops": o, "dense_ops": dense, "bytes": b})
        return {"ops": ops_total, "bytes": bytes_total, "layers": per_layer}


class ShardedFcNetwork:
    """BASELINE config 5: a batch-sharded convolution trunk followed by a block-row-sharded FC.

    Every rank runs the trunk (all layers up to the global average pool) on its own ``batch_per_rank`` images - no
    collective.  The pooled int8 features are written by the pool kernel straight into the rank's rows of the
    ``[world * batch_per_rank, C]`` feature matrix, one in-place all-gather makes it whole on every rank, and the FC runs with
    its block-rows (output channels) split over the ranks: each rank multiplies ALL rows by its own weight slice, the GEMM
    epilogue writes the channel slice into the shard-major logits buffer and a second in-place all-gather completes it
    (``parallel.ShardedBsrLinear``).  ``logits()`` is then a view ``[world * batch_per_rank, num_classes]`` - identical on all
    ranks and bit-equal to the single-GPU network on the concatenated batch."""

    def __init__(self, specs: List[ConvSpec], sparsity_pct: float, batch_per_rank: int, group=None, bias_range: int = 0):
        from . import parallel
        if specs[-1].kind != "fc" or specs[-2].kind != "avgpool":
            raise ValueError("expected a network that ends in avgpool + fc")
        self.rank, self.world = parallel._rank_world(group)
        self.group, self.batch = group, int(batch_per_rank)
        fc_spec = specs[-1]
        self.trunk = BsrNetwork(specs[:-1], sparsity_pct, batch_per_rank, bias_range=bias_range)
        fc_idx = sum(1 for sp in specs[:-1] if sp.kind in ("conv", "fc"))          # same seed as the unsharded network
        syn = synthetic_conv_weights(fc_spec, sparsity_pct, fc_idx, bias_range)
        q, w_scales = exporters.quantize_symmetric_per_channel(syn["w2"], device=True)
        self.fc_bsr = exporters.build_bsr_14x14_int8_direct(q, device=True)
        host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in self.fc_bsr.items()}
        self.fc = parallel.ShardedBsrLinear(host, fc_spec.c_out, group=group)
        self.fc_bias = None if syn["bias"] is None else torch.as_tensor(syn["bias"], dtype=torch.int32).cuda()
        self.fc_spec = fc_spec
        C = specs[-2].c_out
        self.features = torch.zeros((self.world * self.batch, C), dtype=torch.int8, device="cuda")
        # the pool kernel writes this rank's rows of the gathered feature matrix directly
        self.trunk.buffers[specs[-2].name] = self.features[self.rank * self.batch:(self.rank + 1) * self.batch]
        self._buf: Optional[torch.Tensor] = None

    def trunk_forward(self, x: torch.Tensor) -> None:
        self.trunk.forward(x)

    def head_forward(self) -> torch.Tensor:
        """Feature all-gather -> sharded FC -> logits all-gather, all on the current stream.  Returns the logits view."""
        from . import parallel
        mine = self.features[self.rank * self.batch:(self.rank + 1) * self.batch]
        parallel.gather_rows(mine, self.features, self.group)
        self._buf = self.fc.local_gemm(self.features, "i32", bias=self.fc_bias)
        self.fc.gather(self._buf)
        return self.fc.result(self._buf)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.trunk_forward(x)
        return self.head_forward()


class ResNetInference:
    """Python twin of the reference's ``ResNetInference`` engine (hw/sim/cpp/include/resnet_inference.hpp:180-271; its
    ``load_model`` / ``run_inference`` are TODO stubs there) on top of :class:`BsrNetwork`.

    ``load_model(weights_dir)`` reads the directory layout the header documents - ``<layer>_weight_int8.npy`` (OIHW or
    [out, in] int8), ``<layer>_weight_scales.npy`` (float32 per output channel) and optionally ``<layer>_bias_int32.npy`` -
    packs every layer into 14x14 BSR on the GPU (all-zero blocks are dropped, as ``build_bsr_14x14_int8_direct`` does) and
    uploads the plans.  ``run_inference`` takes int8 NCHW images that are already quantised with ``s_in``."""

    def __init__(self, batch: int = 1, image: int = 224, num_classes: int = 1000, s_in: float = S_ACT_IN,
                 s_out: float = S_ACT_OUT):
        self.batch, self.image, self.num_classes = batch, image, num_classes
        self.s_in, self.s_out = s_in, s_out
        self.specs = resnet18_specs(image, num_classes)
        self.net: Optional[BsrNetwork] = None

    def load_model(self, weights_dir: str) -> None:
        import os
        layers: Dict[str, BsrLayer] = {}
        s_in_of = input_scales(self.specs, self.s_in, self.s_out)      # each layer reads at its producer's output scale
        for sp in self.specs:
            if sp.kind not in ("conv", "fc"):
                continue
            w = np.load(os.path.join(weights_dir, f"{sp.name}_weight_int8.npy"))
            if w.dtype != np.int8 or w.shape[0] != sp.c_out or w.size != sp.c_out * sp.K:
                raise ValueError(f"{sp.name}: expected int8 weights of shape [{sp.c_out}, {sp.K}] (or OIHW)")
            scales = np.load(os.path.join(weights_dir, f"{sp.name}_weight_scales.npy")).astype(np.float32).reshape(-1)
            bias_path = os.path.join(weights_dir, f"{sp.name}_bias_int32.npy")
            bias = np.load(bias_path).astype(np.int32) if os.path.exists(bias_path) else None
            bsr = exporters.build_bsr_14x14_int8_direct(torch.from_numpy(w.reshape(sp.c_out, -1)).cuda(), device=True)
            layers[sp.name] = BsrLayer(sp, bsr=bsr, w_scales=scales, bias=bias, s_in=s_in_of[sp.name], s_out=self.s_out,
                                       group_rows=(8 if sp.kind == "fc" else 0))
        self.net = BsrNetwork(self.specs, 0.0, self.batch, layers=layers, s_input=self.s_in, s_out=self.s_out)

    def load_synthetic(self, sparsity_pct: float = 70.0, bias_range: int = 0, chain_scales: bool = False) -> None:
        """Random-init weights of the architecture instead of files (the benchmark recipe of SURVEY.md 8d: He weights, the
        reference's block-mask recipe, per-channel INT8) - there are no checkpoints offline."""
        self.net = BsrNetwork(self.specs, sparsity_pct, self.batch, bias_range=bias_range, s_input=self.s_in, s_out=self.s_out,
                              chain_scales=chain_scales)

    def _require(self) -> BsrNetwork:
        if self.net is None:
            raise RuntimeError("Weights not loaded")          # AcceleratorError::NOT_READY in the C++ twin
        return self.net

    def run_inference_pipelined(self, host_batches, host_logits) -> None:
        """The serving loop for HOST buffers: ``host_batches[i]`` (pinned int8 [batch, 3, H, W]) -> ``host_logits[i]`` (pinned
        int32 [batch, num_classes]) for every i, two batches in flight: two CUDA graphs of the network alternate, each reading
        its own input buffer, so the host-to-device copy of batch i+1 lands in the idle graph's input on a copy stream while
        batch i computes (no staging copy on the device), and the logits are read back on a third stream under the next
        batch's kernels.  Returns when the work is ENQUEUED; synchronise the current stream (or record an event) before
        reading the logits.  Buffers may repeat (a ring of pinned buffers)."""
        net = self._require()
        cur = torch.cuda.current_stream()
        if net.graph is None:
            net.capture(host_batches[0].to("cuda"))
        net.capture_alt()
        if getattr(self, "_pipe", None) is None:
            self._pipe = {"copy": torch.cuda.Stream(), "out": torch.cuda.Stream(),
                          "staged": [torch.cuda.Event(), torch.cuda.Event()], "consumed": [torch.cuda.Event(), torch.cuda.Event()],
                          "computed": torch.cuda.Event(), "read_back": torch.cuda.Event(), "used": [False, False], "any": False}
        P = self._pipe
        inputs, graphs = (net.static_in, net.static_in_alt), (net.graph, net.graph_alt)
        start = torch.cuda.Event()
        start.record(cur)
        P["copy"].wait_event(start)                     # copies of this call start after the work already enqueued
        logits_dev = net.buffers[self.specs[-1].name]
        for i, (xb, yb) in enumerate(zip(host_batches, host_logits)):
            b = i & 1
            with torch.cuda.stream(P["copy"]):
                if P["used"][b]:
                    P["copy"].wait_event(P["consumed"][b])          # the graph that read this input buffer has finished
                inputs[b].copy_(xb, non_blocking=True)               # pinned host -> the graph's own (padded-row) input
                P["staged"][b].record(P["copy"])
            cur.wait_event(P["staged"][b])
            if P["any"]:
                cur.wait_event(P["read_back"])                       # the previous logits have left the (shared) output buffer
            graphs[b].replay()
            P["consumed"][b].record(cur)
            P["computed"].record(cur)
            P["used"][b] = True
            with torch.cuda.stream(P["out"]):                        # the read-back overlaps the next batch's kernels
                P["out"].wait_event(P["computed"])
                yb.copy_(logits_dev, non_blocking=True)
                P["read_back"].record(P["out"])
            P["any"] = True
        cur.wait_event(P["read_back"])                  # "synchronise the current stream" covers the last read-back too

    def run_inference(self, images: torch.Tensor) -> torch.Tensor:
        """int8 [batch, 3, H, W] -> INT32 logits [batch, num_classes] (the FC accumulators, as the reference keeps them)."""
        net = self._require()
        x = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
        x = x.to("cuda", non_blocking=True)
        if net.graph is None:
            net.capture(x)
        return net.replay(x).clone()

    def get_top_k(self, logits: torch.Tensor, k: int = 5):
        w = torch.from_numpy(self._require().layers["fc"].w_scales).to(logits.device)
        s_fc_in = self._require().layers["fc"].s_in                                  # the pooled features' scale
        real = logits.to(torch.float32) * (s_fc_in * w)[: logits.shape[1]]          # de-quantise per channel before ranking
        prob = torch.softmax(real, dim=1)
        p, i = prob.topk(k, dim=1)
        return i.cpu().numpy(), p.cpu().numpy()

    def benchmark(self, num_runs: int = 100) -> Dict:
        net = self._require()
        x = torch.randint(-128, 128, (self.batch, 3, self.image, self.image), dtype=torch.int8, device="cuda")
        if net.graph is None:
            net.capture(x)
        for _ in range(3):
            net.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(num_runs):
            net.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / num_runs
        return {"ms_per_batch": ms, "images_per_s": self.batch / ms * 1e3, "batch": self.batch}

    def get_layer_sparsity(self) -> List[Tuple[str, float]]:
        out = []
        for sp in self.specs:
            if sp.name in self._require().layers:
                lay = self.net.layers[sp.name]
                total = lay.plan.n_block_rows * lay.plan.n_block_cols
                out.append((sp.name, 1.0 - lay.plan.num_blocks / max(total, 1)))
        return out

    def get_model_sparsity(self) -> float:
        net = self._require()
        kept = sum(l.plan.num_blocks for l in net.layers.values())
        total = sum(l.plan.n_block_rows * l.plan.n_block_cols for l in net.layers.values())
        return 1.0 - kept / max(total, 1)

    def print_model_summary(self) -> None:
        for name, s in self.get_layer_sparsity():
            print(f"{name:24s} block sparsity {100 * s:5.1f} %")
        print(f"{'model':24s} block sparsity {100 * self.get_model_sparsity():5.1f} %")
