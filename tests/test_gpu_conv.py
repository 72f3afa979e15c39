"""GPU parity: implicit-im2col BSR convolution (C ABI) vs the oracle and the C++-golden fixtures. Bit-exact."""
import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _plan(bsr):
    from resnet_accel_b200.ops import BsrPlan
    return BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])


def test_cpp_golden_conv_cases(golden):
    """conv2d_int8_im2col outputs recorded from the reference C++ golden (INT32, with bias)."""
    import torch
    c = golden("cpp_golden_cases.npz")
    for i in range(5):
        x, w, bias, ref = c[f"conv{i}_x"], c[f"conv{i}_w"], c[f"conv{i}_bias"], c[f"conv{i}_out"]
        k, s, p = c[f"conv{i}_geom"].tolist()
        bsr = O.build_bsr_14x14_int8_direct(w.reshape(w.shape[0], -1))
        out = _plan(bsr).conv(torch.from_numpy(x[None]).cuda(), k, s, p, w.shape[0], out_kind="i32",
                              bias=bias).cpu().numpy()
        assert np.array_equal(out[0], ref), i


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,s,p,density", [
    (2, 3, 20, 20, 16, 7, 2, 3, 1.0),      # ResNet stem shape family
    (3, 16, 14, 14, 32, 3, 1, 1, 0.5),
    (2, 32, 15, 13, 20, 3, 2, 1, 0.3),
    (5, 64, 7, 7, 64, 3, 1, 1, 0.3),        # 49 positions per image: tiles straddle images
    (2, 64, 9, 9, 128, 1, 2, 0, 0.5),       # 1x1 stride-2 downsample
    (1, 1, 28, 28, 32, 3, 1, 0, 1.0),       # MNIST conv1
    (2, 40, 12, 12, 30, 3, 1, 1, 0.0),      # no stored blocks at all
])
def test_conv_layer_vs_oracle(B, Cin, H, W, Cout, k, s, p, density):
    import torch
    rng = np.random.default_rng(B * 7 + Cin + Cout + k)
    K = Cin * k * k
    Wm = rng.integers(-128, 128, (Cout, K), dtype=np.int8)
    nbr, nbc = -(-Cout // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    bias = rng.integers(-1000, 1000, Cout, dtype=np.int32)
    sf = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32)
    Ho, Wo = O.conv_out_hw(H, W, k, s, p)
    res = rng.integers(-128, 128, (B, Cout, Ho, Wo), dtype=np.int8)
    plan = _plan(bsr)
    xd = torch.from_numpy(x).cuda()
    # int8, relu, bias, residual
    ref, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, k, s, p, bias=bias,
                                       relu=True, sf=sf, residual=res, res_scales=(0.05, 0.02, 0.04))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = plan.conv(xd, k, s, p, Cout, out_kind="i8", chan_scale=sf, bias=bias, relu=True,
                    residual=torch.from_numpy(res).cuda(), res_scales=(0.05, 0.02, 0.04), sat_count=cnt).cpu().numpy()
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)
    assert int(cnt.item()) == sat
    # int32 against the NumPy oracle (independent of the C port)
    ref32, _ = O.conv2d_bsr_layer(x[:1], bsr["indptr"], bsr["indices"], bsr["data"], Cout, k, s, p, bias=bias)
    out32 = plan.conv(xd[:1], k, s, p, Cout, out_kind="i32", bias=bias).cpu().numpy()
    assert np.array_equal(out32, ref32)


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,s,p", [
    (3, 3, 40, 48, 20, 7, 2, 3),        # stem family, rows already 16-byte aligned (dense, W % 16 == 0)
    (2, 30, 28, 28, 40, 3, 1, 1),       # padded rows: TMA tiles, one image per tile
    (6, 16, 7, 7, 32, 3, 1, 1),         # 49 positions per image: one tile spans three images (three TMA boxes)
    (2, 20, 30, 30, 24, 3, 2, 1),       # stride 2
    (1, 64, 56, 56, 64, 3, 1, 1),       # ResNet-18 layer1 geometry
])
def test_conv_padded_rows_tma_path(B, Cin, H, W, Cout, k, s, p):
    """Activations with 16-byte aligned rows (ops.alloc_padded) take the TMA loader; output rows padded as well, and
    the residual read through the same padded layout.  Bit-exact against the oracle, pad bytes stay zero."""
    import torch
    from resnet_accel_b200 import ops
    rng = np.random.default_rng(B + Cin + H + Cout)
    K = Cin * k * k
    Wm = rng.integers(-128, 128, (Cout, K), dtype=np.int8)
    nbr, nbc = -(-Cout // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < 0.4
    Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    bias = rng.integers(-1000, 1000, Cout, dtype=np.int32)
    sf = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32)
    Ho, Wo = O.conv_out_hw(H, W, k, s, p)
    res = rng.integers(-128, 128, (B, Cout, Ho, Wo), dtype=np.int8)
    plan = _plan(bsr)
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    rd = ops.alloc_padded(res.shape)
    rd.copy_(torch.from_numpy(res).cuda())
    out = ops.alloc_padded(res.shape)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    plan.conv(xd, k, s, p, Cout, out_kind="i8", chan_scale=sf, bias=bias, relu=False, residual=rd,
              res_scales=(0.05, 0.05, 0.05), sat_count=cnt, relu_out=True, out=out)
    ref, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, k, s, p, bias=bias,
                                       relu=False, sf=sf, residual=res, res_scales=(0.05, 0.05, 0.05))
    assert np.array_equal(out.cpu().numpy(), np.maximum(ref, 0))
    assert int(cnt.item()) == sat
    base = out._base if out._base is not None else out
    assert int(base[..., Wo:].abs().sum().item()) == 0          # the epilogue never writes the row padding
    ref32, _ = O.conv2d_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, k, s, p, bias=bias)
    assert np.array_equal(plan.conv(xd, k, s, p, Cout, out_kind="i32", bias=bias).cpu().numpy(), ref32)


def test_pools_padded_rows():
    import torch
    from resnet_accel_b200 import ops
    rng = np.random.default_rng(5)
    x = rng.integers(-128, 128, (3, 5, 30, 44), dtype=np.int8)
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    out = ops.alloc_padded((3, 5, 15, 22))
    ops.maxpool_i8(xd, 3, 2, 1, out=out)                          # 3x3/2 pad 1: the vectorised stem pool
    want = np.stack([O.maxpool2d_int8(x[b], 3, 2, 1) for b in range(3)])
    assert np.array_equal(out.cpu().numpy(), want)
    got = ops.avgpool_i8(xd).cpu().numpy()
    assert np.array_equal(got, np.stack([O.avgpool_global_int8(x[b]) for b in range(3)]))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 37, 7, 7), (2, 9, 14, 14), (3, 4, 19, 16), (1, 3, 1, 1), (2, 6, 9, 3)])
def test_avgpool_padded_16_byte_rows(shape):
    """Global average pool over tensors whose rows are padded to 16 bytes (the layer4 output of the network): the padding
    bytes hold junk here and must not count."""
    import torch
    from resnet_accel_b200 import ops
    rng = np.random.default_rng(sum(shape))
    x = rng.integers(-128, 128, shape, dtype=np.int8)
    xd = ops.alloc_padded(shape)
    base = xd._base if xd._base is not None else xd
    base.copy_(torch.randint(-128, 128, base.shape, dtype=torch.int8, device="cuda"))
    xd.copy_(torch.from_numpy(x).cuda())
    got = ops.avgpool_i8(xd).cpu().numpy()
    assert np.array_equal(got, np.stack([O.avgpool_global_int8(x[b]) for b in range(shape[0])]))

