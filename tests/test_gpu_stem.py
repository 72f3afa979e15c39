"""GPU parity: the fused stem kernel (7x7/2 convolution + ReLU/requant + 3x3/2 max-pool, csrc/stem_ws.cuh) against the
oracle's conv2d + maxpool2d_int8.  Bit-exact, including the saturation count over the (never materialised) conv outputs."""
import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _stem_case(B, Cin, H, W, Cout, density, seed, expect_fused=True, bias_on=True):
    import torch
    from resnet_accel_b200 import _lib, ops
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(seed)
    K = Cin * 49
    Wm = rng.integers(-128, 128, (Cout, K), dtype=np.int8)
    nbr, nbc = -(-Cout // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    bias = rng.integers(-2000, 2000, Cout, dtype=np.int32) if bias_on else None
    sf = rng.uniform(2e-4, 3e-3, Cout).astype(np.float32)
    plan = BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    before = _lib.lib().accel_debug_counter(0)
    out = ops.conv_pool(plan, xd, Cout, chan_scale=sf, bias=bias, relu=True, sat_count=cnt)
    torch.cuda.synchronize()
    assert (_lib.lib().accel_debug_counter(0) == before + 1) == expect_fused
    conv, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, 7, 2, 3, bias=bias, relu=True, sf=sf)
    ref = np.stack([O.maxpool2d_int8(conv[b], 3, 2, 1) for b in range(B)])
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert int(cnt.item()) == sat
    base = out._base if out._base is not None else out
    assert int(base[..., ref.shape[-1]:].abs().sum().item()) == 0


@pytest.mark.parametrize("B,Cin,H,W,Cout,density", [
    (3, 3, 64, 64, 64, 0.5),          # odd batch: the last image pair is half empty
    (2, 3, 224, 224, 64, 0.3),        # the ResNet-18 stem
    (4, 3, 32, 96, 20, 0.6),          # few channels: three lane quadrants idle or partial
    (2, 1, 40, 64, 14, 1.0),          # one input channel, dense
    (2, 4, 36, 128, 64, 0.4),         # four input channels
    (2, 3, 16, 32, 30, 0.0),          # no stored blocks
    (1, 3, 4, 32, 8, 1.0),            # batch 1, the smallest image the kernel takes (one pooled row)
    (5, 2, 100, 224, 64, 0.5),        # full-width rows, two channels, odd batch
])
def test_stem_fused_vs_oracle(B, Cin, H, W, Cout, density):
    _stem_case(B, Cin, H, W, Cout, density, seed=B + Cin + H + W + Cout)


@pytest.mark.parametrize("rows", [1, 2, 3, 5, 100])
def test_stem_strip_heights(rows, monkeypatch):
    """An item of the fused kernel is a strip of G pooled rows (chosen on the host for the least work per SM); every strip
    height gives the same bytes and the same saturation count, including strips that do not divide the pooled height."""
    from resnet_accel_b200 import _lib
    monkeypatch.setenv("ACCEL_STEM_ROWS", str(rows))
    _lib.lib().accel_debug_set_timeline(None)          # re-reads the developer switches
    try:
        _stem_case(3, 3, 64, 64, 64, 0.5, seed=21)      # 16 pooled rows, odd batch
        _stem_case(2, 3, 44, 96, 40, 0.5, seed=22)      # 11 pooled rows
    finally:
        monkeypatch.delenv("ACCEL_STEM_ROWS")
        _lib.lib().accel_debug_set_timeline(None)


def test_stem_many_rows_persistent():
    _stem_case(40, 3, 96, 160, 64, 0.3, seed=3)


def test_stem_fallback_geometry():
    """Widths that are not a multiple of 32: the unfused convolution + pool pair runs instead, same result."""
    _stem_case(2, 3, 40, 40, 32, 0.5, seed=4, expect_fused=False)
    _stem_case(2, 3, 64, 64, 80, 0.5, seed=5, expect_fused=False)
