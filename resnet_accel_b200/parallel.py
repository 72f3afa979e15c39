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

from typing import Dict, List, Sequence, Tuple

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


class ShardedBsrLinear:
    """``Y = epilogue(X @ W^T)`` with W's block-rows split over the ranks of ``group`` and one all-gather.

    Each rank builds a plan only for its own block-rows (``resnet_accel_b200.ops.BsrPlan``), computes its channel
    slice with the tcgen05 kernel and contributes it to the gather; ``forward`` returns the full [M, N]."""

    def __init__(self, bsr: Dict, n_out: int, group=None):
        from . import ops
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ranges = shard_block_rows(np.asarray(_host(bsr["indptr"])), self.world)
        self.n_out = int(n_out)
        br0, br1 = self.ranges[self.rank]
        self.br0, self.br1 = br0, br1
        mine = slice_bsr(bsr, br0, br1)
        self.plan = ops.BsrPlan(mine["indptr"], mine["indices"], mine["data"], n_block_cols=bsr["num_block_cols"]) \
            if br1 > br0 else None

    def forward(self, x: torch.Tensor, out_kind: str = "i32", chan_scale=None, bias=None, relu: bool = False) -> torch.Tensor:
        c0, c1 = self.br0 * BLOCK, self.br1 * BLOCK
        dt = {"i8": torch.int8, "i32": torch.int32, "f32": torch.float32}[out_kind]
        if self.plan is not None:
            sl = slice(c0, c1)
            local = self.plan.gemm(x, out_kind, chan_scale=None if chan_scale is None else _pad_to(chan_scale, c1)[sl],
                                   bias=None if bias is None else _pad_to(bias, c1)[sl], relu=relu)
        else:
            local = torch.empty((x.shape[0], 0), dtype=dt, device=x.device)
        return all_gather_channels(local, self.ranges, self.group)[:, :self.n_out]


def _pad_to(v, n: int) -> torch.Tensor:
    t = torch.as_tensor(v)
    if t.numel() >= n:
        return t
    return torch.cat([t, t.new_zeros(n - t.numel())])


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
