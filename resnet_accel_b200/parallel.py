"""Multi-GPU partitioning of the BSR path: one process per GPU, ``torch.distributed``.

Two ways the path shards:

* **batch (M) sharding** - rows of the activation matrix are independent, every rank holds all BSR weights and
  ``1/world`` of the images; **no collective** on the data path (``shard_batch``).  This is how the ResNet-18
  benchmark scales (bench.py).
* **block-row (output-channel) sharding** - in Convention B every output channel belongs to exactly one
  block-row (``sw/golden/golden_fc1_test.py:78-106``), so a rank that owns block-rows ``[br0, br1)`` computes
  ``Y[:, 14*br0 : 14*br1]`` on its own; one all-gather along the channel axis rebuilds ``Y``
  (``shard_block_rows`` / ``slice_bsr`` / ``all_gather_channels``).  Natural only for wide, short layers
  (the FC of ResNet-50: 72 block-rows, M = batch).

The helpers are pure host logic + ``torch.distributed`` calls, so they run on the ``gloo`` backend on CPU
(tests/test_parallel_gloo.py) exactly as they do on NCCL.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

BLOCK = 14


def shard_batch(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``n`` images for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_block_rows(row_ptr: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split block-rows into ``world`` contiguous ranges balanced by stored blocks (not by row count).

    Greedy on the prefix sums of ``row_ptr``: range g ends at the first block-row whose cumulative block count
    reaches ``(g+1)/world`` of the total; every range may be empty only if there are fewer block-rows than ranks."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    nbr = len(rp) - 1
    total = int(rp[-1])
    cuts = [0]
    for g in range(1, world):
        if total == 0:
            c = (nbr * g) // world
        else:
            c = int(np.searchsorted(rp, total * g / world, side="left"))
        c = min(max(c, cuts[-1]), nbr)
        cuts.append(c)
    cuts.append(nbr)
    return [(cuts[g], cuts[g + 1]) for g in range(world)]


def slice_bsr(bsr: Dict, br0: int, br1: int) -> Dict:
    """The BSR sub-matrix of block-rows [br0, br1) in the exporter's dict layout (export_bsr_14x14.py:206-235)."""
    rp = np.asarray(bsr["indptr"], dtype=np.int64)
    b0, b1 = int(rp[br0]), int(rp[br1])
    data, idx = bsr["data"][b0:b1], bsr["indices"][b0:b1]
    out = dict(bsr)
    out.update({
        "data": data, "indices": idx,
        "indptr": (rp[br0:br1 + 1] - b0).astype(np.int32) if not isinstance(bsr["indptr"], torch.Tensor)
        else (bsr["indptr"][br0:br1 + 1] - b0),
        "num_blocks": b1 - b0, "num_block_rows": br1 - br0,
    })
    return out


def all_gather_channels(local: torch.Tensor, ranges: Sequence[Tuple[int, int]], group=None) -> torch.Tensor:
    """All-gather ``local`` [M, 14*(br1-br0)] of every rank along the channel axis -> [M, 14*nbr].

    Shards may differ in width, so every rank pads its columns to the widest shard, one
    ``all_gather_into_tensor`` moves the [world, M, wmax] buffer (NCCL over NVLink on GPUs), and the
    valid columns are copied out.  The payload is small (ResNet-50 FC at batch 1024: ~1 MB in total)."""
    world = dist.get_world_size(group)
    widths = [(b1 - b0) * BLOCK for b0, b1 in ranges]
    wmax = max(widths) if widths else 0
    M = local.shape[0]
    send = local.new_zeros((M, wmax))
    send[:, :local.shape[1]] = local
    recv2 = local.new_empty((world * M, wmax))                  # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(recv2, send.contiguous(), group=group)
    recv = recv2.view(world, M, wmax)
    out = local.new_empty((M, sum(widths)))
    col = 0
    for g, w in enumerate(widths):
        out[:, col:col + w] = recv[g, :, :w]
        col += w
    return out


def shard_block_rows_even(n_block_rows: int, world: int) -> List[Tuple[int, int]]:
    """Equal block-row counts per rank (``n_block_rows`` must divide): every shard then has the same width, so the gathered
    buffer needs no padding and the full result is a plain view of it (ResNet-50 FC: 72 block-rows -> 9 per rank on 8)."""
    if n_block_rows % world:
        raise ValueError(f"{n_block_rows} block-rows do not divide over {world} ranks")
    per = n_block_rows // world
    return [(g * per, (g + 1) * per) for g in range(world)]


class ShardedBsrLinear:
    """``Y = epilogue(X @ W^T)`` with W's block-rows split over the ranks of ``group`` and ONE all-gather.

    Each rank builds a plan only for its own block-rows (``ops.BsrPlan``).  The result lives channel-major: the gather
    buffer is ``Yt[world, wmax, M]`` and the epilogue of the rank's GEMM writes its ``[width, M]`` slice straight into
    ``Yt[rank]`` (``BsrPlan.gemm(out_t=...)``: no staging tensor, no padding pass).  ``all_gather_into_tensor`` then runs in
    place on that buffer (the send buffer IS the rank's slice of the receive buffer), on the caller's stream or on the side
    stream of ``forward_async`` so that it overlaps the next batch's convolutions.  With equal shards (the block-rows divide
    over the ranks) the full ``[M, N]`` result is a strided VIEW of the buffer - nothing is copied after the collective;
    unequal shards (balanced by stored blocks) cost one column copy per shard.

    The activations ``x`` must be the same ``[M, K]`` matrix on every rank (after a batch-sharded trunk: one all-gather of the
    int8 features, ``gather_rows``)."""

    def __init__(self, bsr: Dict, n_out: int, group=None, balance: str = "auto", build_plan: bool = True,
                 rank: Optional[int] = None, world: Optional[int] = None):
        self.group = group
        self.rank, self.world = _rank_world(group)
        if rank is not None and world is not None:      # explicit placement (single-process emulation of the ranks in tests)
            self.rank, self.world = int(rank), int(world)
        nbr = len(_host(bsr["indptr"])) - 1
        if balance == "rows" or (balance == "auto" and nbr % self.world == 0):
            self.ranges = shard_block_rows_even(nbr, self.world)
        else:
            self.ranges = shard_block_rows(np.asarray(_host(bsr["indptr"])), self.world)
        self.widths = [(b1 - b0) * BLOCK for b0, b1 in self.ranges]
        self.wmax = max(self.widths) if self.widths else 0
        self.even = len(set(self.widths)) <= 1
        self.n_out = int(n_out)
        br0, br1 = self.ranges[self.rank]
        self.br0, self.br1 = br0, br1
        self.mine = slice_bsr(bsr, br0, br1)
        self.plan = None
        if build_plan and br1 > br0:
            from . import ops
            self.plan = ops.BsrPlan(self.mine["indptr"], self.mine["indices"], self.mine["data"], n_block_cols=bsr["num_block_cols"])
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self._comm_stream = None

    def buffer(self, M: int, dtype: torch.dtype, device, slot: int = 0) -> torch.Tensor:
        """The gather buffer ``[world, wmax, M]`` for this (M, dtype); ``slot`` selects one of several in-flight buffers."""
        key = (int(M), dtype, str(device), int(slot))
        if key not in self._bufs:
            self._bufs[key] = torch.zeros((self.world, self.wmax, int(M)), dtype=dtype, device=device)
        return self._bufs[key]

    def local_gemm(self, x: torch.Tensor, out_kind: str = "i32", chan_scale=None, bias=None, relu: bool = False,
                   slot: int = 0) -> torch.Tensor:
        """This rank's channel slice, written by the GEMM epilogue into its place in the gather buffer.  Returns the buffer."""
        c0, c1 = self.br0 * BLOCK, self.br1 * BLOCK
        dt = {"i8": torch.int8, "i32": torch.int32, "f32": torch.float32}[out_kind]
        buf = self.buffer(x.shape[0], dt, x.device, slot)
        if self.plan is not None:
            sl = slice(c0, c1)
            self.plan.gemm(x, out_kind, chan_scale=None if chan_scale is None else _pad_to(chan_scale, c1)[sl],
                           bias=None if bias is None else _pad_to(bias, c1)[sl], relu=relu,
                           out_t=buf[self.rank, : c1 - c0])
        return buf

    def gather(self, buf: torch.Tensor) -> None:
        """In-place all-gather of the shard-major buffer on the current stream."""
        if self.world == 1:
            return
        flat = buf.view(self.world * self.wmax, buf.shape[2])
        dist.all_gather_into_tensor(flat, buf[self.rank], group=self.group)

    def result(self, buf: torch.Tensor) -> torch.Tensor:
        """``[M, n_out]`` from the gathered buffer: a view when the shards are equal, else one column copy per shard."""
        M = buf.shape[2]
        if self.even:
            return buf.view(self.world * self.wmax, M).t()[:, : self.n_out]
        out = buf.new_empty((M, sum(self.widths)))
        col = 0
        for g, w in enumerate(self.widths):
            out[:, col:col + w] = buf[g, :w].t()
            col += w
        return out[:, : self.n_out]

    def forward(self, x: torch.Tensor, out_kind: str = "i32", chan_scale=None, bias=None, relu: bool = False) -> torch.Tensor:
        buf = self.local_gemm(x, out_kind, chan_scale, bias, relu)
        self.gather(buf)
        return self.result(buf)

    def forward_async(self, x: torch.Tensor, out_kind: str = "i32", chan_scale=None, bias=None, relu: bool = False,
                      slot: int = 0):
        """GEMM on the current stream, the all-gather on a side stream: returns ``(buffer, event)``; the caller keeps
        launching the next batch on the current stream and waits on ``event`` before it reads ``result(buffer)``."""
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        buf = self.local_gemm(x, out_kind, chan_scale, bias, relu, slot=slot)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        done = torch.cuda.Event()
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ready)
            self.gather(buf)
            done.record(self._comm_stream)
        return buf, done


def gather_rows(local: torch.Tensor, out: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of equal row blocks: ``local`` [m, K] of every rank -> ``out`` [world * m, K] (rank-major).  ``local`` may be
    the rank's own slice of ``out`` (in-place collective: the producer writes its rows where they belong)."""
    if _rank_world(group)[1] > 1:
        dist.all_gather_into_tensor(out, local, group=group)
    elif out.data_ptr() != local.data_ptr():
        out.copy_(local)
    return out


def _rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _pad_to(v, n: int) -> torch.Tensor:
    t = torch.as_tensor(v)
    if t.numel() >= n:
        return t
    return torch.cat([t, t.new_zeros(n - t.numel())])


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
