"""BASELINE config 2: synthetic BSR INT8 GEMM 4096 x 4096 x 4096, block-sparsity sweep 0 / 50 / 70 / 90 %.

Inputs follow SURVEY.md 8d: ``default_rng(0)`` weights and activations, the reference's block mask recipe
(``create_sparse_mask(shape, pct, block_size=14, seed=42)``, sw/exporters/export_transformer.py:19-60), packed by
``build_bsr_14x14_int8_direct`` (sw/training/export_bsr_14x14.py:406-484); output INT32 [4096, 4102] as
``gemm_bsr_int8_golden`` returns it (sw/golden/golden_fc1_test.py:49-108).

Parity at full size: (1) 256 sampled rows against the plain-C oracle, bit for bit; (2) the whole matrix across two
independent schedules - the dense-equivalent CTA-pair kernel (csrc/gemm_ws.cuh) against the block-gather kernel
(csrc/bsr_tcp.cuh, block-row groups of 8) - equal element for element; (3) the padding columns 4096..4101 are zero.
"""
import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu
N = 4096


@pytest.mark.parametrize("pct", [0.0, 50.0, 70.0, 90.0])
def test_gemm4096_sweep(pct):
    import torch
    from resnet_accel_b200 import _lib, exporters as E, ops
    rng = np.random.default_rng(0)
    W = rng.integers(-128, 128, (N, N), dtype=np.int8)
    mask = E.create_sparse_mask((N, N), pct, block_size=14, seed=42)
    assert np.array_equal(mask, O.create_sparse_mask((N, N), pct, 14, 42))
    W = (W * mask.astype(np.int8)).astype(np.int8)
    A = rng.integers(-128, 128, (N, N), dtype=np.int8)
    bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(W).cuda(), device=True)          # GPU packer
    nbr = nbc = -(-N // 14)
    assert bsr["num_blocks"] == nbr * nbc - int(nbr * nbc * pct / 100)
    rp, ci, blk = (bsr[k].cpu().numpy() for k in ("indptr", "indices", "data"))
    x = torch.from_numpy(A).cuda()
    L = _lib.lib()
    # schedule 1: dense-equivalent kernel, CTA pairs
    plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=nbc)
    before = L.accel_debug_counter(1)
    out = plan.gemm(x, "i32")
    assert L.accel_debug_counter(1) == before + 1, "gemm_ws_kernel did not run"
    assert tuple(out.shape) == (N, nbr * 14) and out.dtype == torch.int32
    # schedule 2: block-gather kernel, block-row groups of 8
    plan2 = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=nbc, group_rows=8)
    plan2.use_gemm_ws = False
    before = L.accel_debug_counter(1)
    out2 = plan2.gemm(x, "i32")
    assert L.accel_debug_counter(1) == before, "the gather kernel was expected"
    assert torch.equal(out, out2)
    assert int(out.to(torch.int64).sum().item()) == int(out2.to(torch.int64).sum().item())
    assert not out[:, N:].any().item()
    # oracle on sampled rows
    rows = np.sort(np.random.default_rng(1).choice(N, 256, replace=False))
    ref = c_oracle.bsr_gemm_i32(A[rows], rp, ci, blk)
    got = out[torch.from_numpy(rows).cuda()].cpu().numpy()
    assert np.array_equal(got, ref)
    # fused epilogue at full size: ReLU + per-channel requant to int8 with the saturation counter, both schedules
    sf = np.random.default_rng(2).uniform(2e-5, 2e-4, nbr * 14).astype(np.float32)
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    q1 = plan.gemm(x, "i8", chan_scale=sf, relu=True, sat_count=cnt[0:1])
    q2 = plan2.gemm(x, "i8", chan_scale=sf, relu=True, sat_count=cnt[1:2])
    assert torch.equal(q1, q2) and int(cnt[0].item()) == int(cnt[1].item())
    want, sat = O.requantize_int32_to_int8(np.maximum(ref, 0), sf[None, :])
    assert np.array_equal(q1[torch.from_numpy(rows).cuda()].cpu().numpy(), want)
