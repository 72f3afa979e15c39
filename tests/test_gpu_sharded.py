"""BASELINE config 5: block-row-sharded FC + all-gather, ResNet-50 topology.

* one GPU: the ranks of the sharded FC are emulated one after the other in one process - every "rank" runs its GEMM, whose
  epilogue writes the channel slice straight into its place of the shard-major buffer; the assembled view must equal the
  unsharded GEMM bit for bit (this is the part of the path the round-end single-GPU run can see);
* two or more GPUs: one process per GPU, NCCL, ``ShardedFcNetwork`` on a down-scaled ResNet-50 - batch-sharded trunk, feature
  all-gather, sharded FC, logits all-gather - against the single-GPU network on the concatenated batch.
"""
import os
import sys

import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.parametrize("world,M", [(2, 1024), (8, 1024), (4, 100), (3, 256)])
def test_sharded_fc_epilogue_writes_gather_buffer(world, M):
    """ResNet-50 FC shape (2048 -> 1000, 72 block-rows) at 70 % block sparsity."""
    import torch
    from resnet_accel_b200 import exporters as E, ops, parallel as P
    rng = np.random.default_rng(world * 10 + M)
    N, K = 1000, 2048
    W = rng.integers(-128, 128, (N, K), dtype=np.int8)
    W = (W * E.create_sparse_mask((N, K), 70.0, block_size=14, seed=42).astype(np.int8)).astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(W)
    A = rng.integers(-128, 128, (M, K), dtype=np.int8)
    bias = rng.integers(-1000, 1000, N, dtype=np.int32)
    x = torch.from_numpy(A).cuda()
    full = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    want = full.gemm(x, "i32", n_channels=N, bias=bias)
    ranks = [P.ShardedBsrLinear(bsr, N, rank=r, world=world) for r in range(world)]
    assert ranks[0].even == (72 % world == 0)
    shared = ranks[0].buffer(M, torch.int32, x.device)
    for r in ranks:                      # what the in-place all-gather leaves on every rank: all slices in one buffer
        r._bufs = ranks[0]._bufs
        buf = r.local_gemm(x, "i32", bias=bias)
        assert buf.data_ptr() == shared.data_ptr()
    got = ranks[0].result(shared)
    assert torch.equal(got, want)
    if ranks[0].even:                    # equal shards: the result is a VIEW of the gather buffer, nothing was copied
        assert got.data_ptr() == shared.data_ptr() and got.stride() == (1, M)
    ref = c_oracle.bsr_gemm_i32(A[:64], bsr["indptr"], bsr["indices"], bsr["data"])[:, :N] + bias[None, :]
    assert np.array_equal(got[:64].cpu().numpy(), ref)


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from resnet_accel_b200 import layers as L
        B = 2
        specs = L.resnet50_specs(image=64, num_classes=1000)
        net = L.ShardedFcNetwork(specs, 70.0, B, bias_range=100)
        x_all = torch.from_numpy(np.random.default_rng(5).integers(-128, 128, (world * B, 3, 64, 64), dtype=np.int8)).cuda()
        logits = net.forward(x_all[rank * B:(rank + 1) * B]).contiguous()
        torch.cuda.synchronize()
        # single-GPU reference on the concatenated batch (same seeds -> same weights)
        ref_net = L.BsrNetwork(specs, 70.0, world * B, bias_range=100)
        ref = ref_net.forward(x_all)
        ok = bool(torch.equal(logits, ref))
        gathered = [torch.empty_like(logits) for _ in range(world)]
        dist.all_gather(gathered, logits)
        same = all(bool(torch.equal(g, logits)) for g in gathered)
        q.put((rank, ok, same, int(logits.abs().sum().item())))
    finally:
        dist.destroy_process_group()


def test_resnet50_sharded_fc_nccl():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, ok, same, checksum in res:
        assert ok and same and checksum > 0, (rank, ok, same, checksum)


def test_resnet50_network_single_gpu_matches_port():
    """The ResNet-50 layer table (torchvision v1.5 topology) through BsrNetwork at 64x64: every convolution
    against the plain-C port, so that the 1x1 / bottleneck / stride-2-on-the-3x3 paths are pinned too."""
    import torch
    from resnet_accel_b200 import layers as L
    from test_gpu_network import _check_layers_against_port
    B = 2
    specs = L.resnet50_specs(image=64, num_classes=1000)
    assert sum(1 for s in specs if s.kind == "conv") == 53 and specs[-1].c_in == 2048
    net = L.BsrNetwork(specs, 70.0, B, bias_range=100)
    x = np.random.default_rng(7).integers(-128, 128, (B, 3, 64, 64), dtype=np.int8)
    s_in = {sp.name: L.S_ACT_IN for sp in specs}
    res = {sp.name: (L.S_ACT_OUT, L.S_ACT_OUT, L.S_ACT_OUT) for sp in specs}
    _check_layers_against_port(net, specs, x, s_in, res)
