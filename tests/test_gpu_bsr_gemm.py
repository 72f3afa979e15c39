"""GPU parity: tcgen05 BSR GEMM (through the C ABI) vs the CPU oracle and the golden fixtures. Bit-exact."""
import hashlib

import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _plan(bsr, **kw):
    from resnet_accel_b200.ops import BsrPlan
    return BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"], **kw)


def _rand_bsr(rng, N, K, density):
    W = rng.integers(-128, 128, (N, K), dtype=np.int8)
    nbr, nbc = -(-N // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    W = W * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:N, :K].astype(np.int8)
    return W, O.build_bsr_14x14_int8_direct(W)


def test_device_check():
    from resnet_accel_b200 import _lib
    assert _lib.lib().accel_device_check() == 0
    assert b"sm_100a" in _lib.lib().accel_version()


def test_fc1_known_answer(golden):
    """SURVEY.md A.5 through the CUDA path."""
    torch = _torch()
    mn, ka = golden("mnist_int8.npz"), golden("fc1_known_answer.npz")
    bsr = O.build_bsr_14x14_int8_direct(mn["fc1_weight_int8"])
    plan = _plan(bsr)
    out = plan.gemm(torch.from_numpy(ka["activations"]).cuda()).cpu().numpy()
    assert out.shape == (1, 140)
    assert np.array_equal(out, ka["output"])
    assert hashlib.sha256(out.tobytes()).hexdigest()[:16] == "82c5a90b98005a39"


def test_golden_i32_cases(golden):
    torch = _torch()
    g = golden("golden_i32_cases.npz")
    from resnet_accel_b200.ops import BsrPlan
    for i in range(6):
        rp, ci, blk = g[f"c{i}_row_ptr"], g[f"c{i}_col_idx"], g[f"c{i}_blocks"]
        nbc = -(-g[f"c{i}_W"].shape[1] // 14)
        plan = BsrPlan(rp, ci, blk, n_block_cols=nbc)
        for kv in ("padK", "rawK"):
            A, Cref = g[f"c{i}_{kv}_A"], g[f"c{i}_{kv}_C"]
            out = plan.gemm(torch.from_numpy(A).cuda()).cpu().numpy()
            assert np.array_equal(out, Cref), (i, kv)
    for tag in ("max", "min", "mixed"):
        bsr = O.build_bsr_14x14_int8_direct(g[f"x_{tag}_W"])
        out = _plan(bsr).gemm(torch.from_numpy(g[f"x_{tag}_A"]).cuda()).cpu().numpy()
        assert np.array_equal(out, g[f"x_{tag}_C"]), tag


@pytest.mark.parametrize("M,N,K,density,group_rows", [
    (1, 14, 14, 1.0, 0), (128, 28, 28, 1.0, 0), (129, 140, 300, 0.5, 0), (257, 500, 1000, 0.3, 0),
    (300, 1000, 512, 0.7, 8), (64, 128, 9216, 0.1, 0), (77, 449, 3000, 0.0, 0), (512, 448, 2304, 1.0, 5),
    (33, 15, 225, 0.6, 0), (200, 600, 237, 0.9, 32),
])
def test_random_vs_oracle(M, N, K, density, group_rows):
    torch = _torch()
    rng = np.random.default_rng(M * 1000 + N + K)
    W, bsr = _rand_bsr(rng, N, K, density)
    A = rng.integers(-128, 128, (M, K), dtype=np.int8)
    ref = c_oracle.bsr_gemm_i32(A, bsr["indptr"], bsr["indices"], bsr["data"])
    plan = _plan(bsr, group_rows=group_rows)
    out = plan.gemm(torch.from_numpy(A).cuda()).cpu().numpy()
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)
    # dense cross-check where the matrix is small enough
    if M * N * K < 2e8:
        dense = A.astype(np.int64) @ W.T.astype(np.int64)
        assert np.array_equal(out[:, :N], dense.astype(np.int32))


def test_odd_alignment_and_strided_rows():
    """lda > K, odd base pointer / odd lda (byte-load path of the producer)."""
    torch = _torch()
    rng = np.random.default_rng(5)
    W, bsr = _rand_bsr(rng, 70, 131, 0.6)
    plan = _plan(bsr)
    buf = torch.from_numpy(rng.integers(-128, 128, (50, 200), dtype=np.int8)).cuda()
    for off, K in ((0, 131), (1, 131), (3, 140)):
        x = buf[:, off:off + K]
        ref = c_oracle.bsr_gemm_i32(x.cpu().numpy(), bsr["indptr"], bsr["indices"], bsr["data"])
        assert np.array_equal(plan.gemm(x).cpu().numpy(), ref), (off, K)


def test_fused_epilogue_requant_relu_bias_residual():
    torch = _torch()
    rng = np.random.default_rng(11)
    M, N, K = 300, 130, 420
    W, bsr = _rand_bsr(rng, N, K, 0.5)
    A = rng.integers(-128, 128, (M, K), dtype=np.int8)
    bias = rng.integers(-2000, 2000, N, dtype=np.int32)
    sf = rng.uniform(2e-4, 3e-3, N).astype(np.float32)
    res = rng.integers(-128, 128, (M, N), dtype=np.int8)
    plan = _plan(bsr)
    x = torch.from_numpy(A).cuda()
    # int32 + bias + relu
    ref32, _ = O.linear_bsr_layer(A, bsr["indptr"], bsr["indices"], bsr["data"], N, bias=bias, relu=True)
    out32 = plan.gemm(x, "i32", n_channels=N, bias=torch.from_numpy(bias).cuda(), relu=True).cpu().numpy()
    assert np.array_equal(out32, ref32)
    # int8 requant, saturation counter, per-channel abs-max
    ref8, sat = O.linear_bsr_layer(A, bsr["indptr"], bsr["indices"], bsr["data"], N, bias=bias, relu=False,
                                   scale_factor=sf)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    amax = torch.zeros(N, dtype=torch.int32, device="cuda")
    out8 = plan.gemm(x, "i8", n_channels=N, chan_scale=sf, bias=bias, sat_count=cnt, chan_absmax=amax).cpu().numpy()
    assert np.array_equal(out8, ref8)
    assert int(cnt.item()) == sat and sat > 0
    acc, _ = O.linear_bsr_layer(A, bsr["indptr"], bsr["indices"], bsr["data"], N, bias=bias)
    assert np.array_equal(amax.cpu().numpy(), np.abs(acc.astype(np.int64)).max(axis=0).astype(np.int32))
    # residual add on top
    ref_r = O.add_residual_int8(ref8, res, 0.05, 0.03, 0.04)
    out_r = plan.gemm(x, "i8", n_channels=N, chan_scale=sf, bias=bias, residual=torch.from_numpy(res).cuda(),
                      res_scales=(0.05, 0.03, 0.04)).cpu().numpy()
    assert np.array_equal(out_r, ref_r)
    # float32 de-quantised logits
    outf = plan.gemm(x, "f32", n_channels=N, chan_scale=sf).cpu().numpy()
    base, _ = O.linear_bsr_layer(A, bsr["indptr"], bsr["indices"], bsr["data"], N)
    assert np.array_equal(outf, base.astype(np.float32) * sf[None, :])


def test_generic_block_kernels_and_convention_a(golden):
    torch = _torch()
    from resnet_accel_b200.ops import bsr_gemm_generic
    rng = np.random.default_rng(2)
    for bh, bw in ((8, 8), (4, 4), (16, 16), (4, 8), (14, 14)):
        W = rng.integers(-128, 128, (50, 77), dtype=np.int8)
        W[: bh * 2] = 0
        bsr = O.build_bsr_from_dense(W, bh, bw)
        A = rng.integers(-128, 128, (9, 77), dtype=np.int8)
        ref = O.bsr_gemm_i32(A, bsr["indptr"], bsr["indices"], bsr["data"])
        out = bsr_gemm_generic(A, bsr["indptr"], bsr["indices"], bsr["data"], ref.shape[1]).cpu().numpy()
        assert np.array_equal(out, ref), (bh, bw)
    c = golden("cpp_golden_cases.npz")
    out = bsr_gemm_generic(c["convA_A"], c["convA_row_ptr"].astype(np.int32), c["convA_col_idx"].astype(np.int32),
                           c["convA_data"].reshape(-1, 14, 14), 45, orient=1).cpu().numpy()
    assert np.array_equal(out, c["convA_C"])


def test_invalid_inputs_raise():
    torch = _torch()
    from resnet_accel_b200 import AcceleratorError
    from resnet_accel_b200.ops import BsrPlan
    blk = np.zeros((2, 14, 14), np.int8)
    with pytest.raises(AcceleratorError):
        BsrPlan([0, 2], [1, 0], blk, n_block_cols=2)          # unsorted col_idx
    with pytest.raises(AcceleratorError):
        BsrPlan([0, 2], [0, 5], blk, n_block_cols=2)          # col out of range
    with pytest.raises(AcceleratorError):
        BsrPlan([1, 2], [0], blk[:1], n_block_cols=2)         # row_ptr[0] != 0
    plan = BsrPlan([0, 1], [0], blk[:1], n_block_cols=1)
    with pytest.raises(AcceleratorError):
        plan.gemm(torch.zeros((4, 14), dtype=torch.int32, device="cuda"))
    with pytest.raises(AcceleratorError):
        plan.gemm(torch.zeros((4, 14), dtype=torch.int8, device="cuda"), "i8")   # chan_scale missing
