"""CPU, world_size 2, gloo: the host side of the multi-GPU path.

Batch sharding needs no collective; block-row sharding needs one all-gather along the channel axis.  The ranks run
the CPU oracle on their shard (tests may use the oracle) and the gathered result must equal the unsharded one."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bsr_oracle as O
from resnet_accel_b200 import parallel as P


def _make(seed=3, N=200, K=300, density=0.4):
    rng = np.random.default_rng(seed)
    W = rng.integers(-128, 128, (N, K), dtype=np.int8)
    nbr, nbc = -(-N // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    keep[3] = False                                           # an empty block-row
    W = W * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:N, :K].astype(np.int8)
    A = rng.integers(-128, 128, (37, K), dtype=np.int8)
    return O.build_bsr_14x14_int8_direct(W), A, N


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bsr, A, N = _make()
        full = O.bsr_gemm_i32(A, bsr["indptr"], bsr["indices"], bsr["data"])
        # --- block-row sharding + all-gather
        ranges = P.shard_block_rows(bsr["indptr"], world)
        br0, br1 = ranges[rank]
        mine = P.slice_bsr(bsr, br0, br1)
        local = O.bsr_gemm_i32(A, mine["indptr"], mine["indices"], mine["data"]) if br1 > br0 else np.zeros((A.shape[0], 0), np.int32)
        got = P.all_gather_channels(torch.from_numpy(np.ascontiguousarray(local)), ranges).numpy()
        ok_rows = np.array_equal(got, full)
        # --- the direct-write layout of ShardedBsrLinear: the rank's slice is written TRANSPOSED into its place of the
        #     shard-major buffer, the all-gather runs in place on it, the result is assembled from (or is a view of) it
        for balance in ("auto", "rows") if (len(bsr["indptr"]) - 1) % world == 0 else ("auto",):
            lin = P.ShardedBsrLinear(bsr, N, balance=balance, build_plan=False)
            buf = lin.buffer(A.shape[0], torch.int32, "cpu")
            b0, b1 = lin.ranges[rank]
            if b1 > b0:
                loc = O.bsr_gemm_i32(A, lin.mine["indptr"], lin.mine["indices"], lin.mine["data"])
                buf[rank, : loc.shape[1]] = torch.from_numpy(np.ascontiguousarray(loc.T))
            lin.gather(buf)
            ok_rows = ok_rows and np.array_equal(lin.result(buf).numpy(), full[:, :N])
        # --- batch sharding: no collective on the data path; gather only to check
        lo, hi = P.shard_batch(A.shape[0], rank, world)
        part = O.bsr_gemm_i32(A[lo:hi], bsr["indptr"], bsr["indices"], bsr["data"])
        parts = [None] * world
        dist.all_gather_object(parts, part)
        ok_batch = np.array_equal(np.concatenate(parts, 0), full)
        q.put((rank, bool(ok_rows), bool(ok_batch), ranges))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_rows, ok_batch, ranges in res:
        assert ok_rows and ok_batch, (rank, ranges)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_block_row_ranges_balance_blocks(world):
    bsr, _, _ = _make(seed=world, N=1000, K=2048, density=0.3)
    rp = np.asarray(bsr["indptr"])
    ranges = P.shard_block_rows(rp, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(rp) - 1
    assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    counts = [int(rp[b1] - rp[b0]) for b0, b1 in ranges]
    assert sum(counts) == int(rp[-1])
    assert max(counts) - min(counts) <= 2 * int(np.diff(rp).max())        # within two block-rows of even
    for n in (0, 1, 7, 256, 1000):
        parts = [P.shard_batch(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1
