"""GPU parity: the weight-stationary 3x3 / stride-1 convolution kernel (csrc/conv_ws.cuh) vs the C oracle. Bit-exact.

The kernel is taken when activations have 16-byte aligned rows (ops.alloc_padded), Cin % 32 == 0 and W <= 62; every case
checks that it really ran (accel_debug_counter) so a silent fall-back to the gather kernels cannot hide a failure."""
import numpy as np
import pytest

from oracle import bsr_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _case(B, Cin, H, W, Cout, density, seed, residual=True, relu=False, relu_out=True, bias_on=True,
          res_scales=(0.05, 0.02, 0.04), wmax=128, sf_fn=None, x_fill=None, bias_mag=1000, expect_fast=None):
    import torch
    from resnet_accel_b200 import _lib, ops
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(seed)
    K = Cin * 9
    Wm = rng.integers(-wmax, wmax, (Cout, K)).astype(np.int8)
    nbr, nbc = -(-Cout // 14), -(-K // 14)
    keep = rng.random((nbr, nbc)) < density
    Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    if x_fill is not None:
        x[B // 2:] = x_fill                  # half of the images constant: the accumulators reach their bound
    bias = rng.integers(-bias_mag, bias_mag, Cout, dtype=np.int32) if bias_on else None
    sf = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32) if sf_fn is None else sf_fn(rng, Cout).astype(np.float32)
    res = rng.integers(-128, 128, (B, Cout, H, W), dtype=np.int8) if residual else None
    plan = BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    # garbage in the row padding of the INPUT must not matter: TMA zero-fills beyond W
    base_in = xd._base if xd._base is not None else xd
    if base_in.shape[-1] > W:
        base_in[..., W:] = 77
    rd = None
    if residual:
        rd = ops.alloc_padded(res.shape)
        rd.copy_(torch.from_numpy(res).cuda())
    out = ops.alloc_padded((B, Cout, H, W))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    before, before_fast = _lib.lib().accel_debug_counter(0), _lib.lib().accel_debug_counter(2)
    plan.conv(xd, 3, 1, 1, Cout, out_kind="i8", chan_scale=sf, bias=bias, relu=relu, residual=rd,
              res_scales=res_scales, sat_count=cnt, relu_out=relu_out, out=out)
    torch.cuda.synchronize()
    assert _lib.lib().accel_debug_counter(0) == before + 1, "the weight-stationary kernel did not run"
    if expect_fast is not None:
        assert _lib.lib().accel_debug_counter(2) - before_fast == int(expect_fast), "wrong epilogue variant"
    ref, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, 3, 1, 1, bias=bias,
                                       relu=relu, sf=sf, residual=res, res_scales=res_scales)
    if relu_out:
        ref = np.maximum(ref, 0)
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert int(cnt.item()) == sat
    base = out._base if out._base is not None else out
    assert int(base[..., W:].abs().sum().item()) == 0          # row padding of the output stays zero


@pytest.mark.parametrize("B,Cin,H,W,Cout,density", [
    (2, 64, 56, 56, 64, 0.3),        # ResNet-18 layer1: 64-byte rows, 2 output rows per tile, weights resident
    (3, 128, 28, 28, 128, 0.3),      # layer2: 32-byte rows, 4 rows per tile, weights resident (4 chunks)
    (3, 256, 14, 14, 256, 0.3),      # layer3: 16-byte rows, 7 rows per tile, two channel groups, weights streamed
    (5, 512, 7, 7, 512, 0.3),        # layer4: one tile per image, four channel groups, 16 chunks streamed
    (2, 32, 9, 20, 40, 0.5),         # ragged last tile (9 rows in tiles of 4), Cout not a multiple of 14
    (2, 64, 13, 37, 200, 0.4),       # two channel groups, the second partial; odd sizes
    (1, 96, 5, 3, 10, 1.0),          # tiny image, dense weights
    (2, 64, 30, 62, 70, 0.2),        # widest supported row (62 + 2 padding pixels)
    (2, 64, 12, 12, 30, 0.0),        # no stored blocks at all
    (2, 160, 11, 14, 129, 0.05),     # very sparse: some (chunk, tap) weight tiles are empty and skipped
    (4, 64, 7, 7, 128, 0.3),         # 7-pixel rows, > 64 channels: two images per staged row ("twin" tiles), weights resident
    (3, 288, 6, 5, 130, 0.4),        # twin tiles, odd batch (the last pair is half empty), streamed weights, two groups
    (1, 32, 1, 1, 16, 1.0),          # a single pixel
    (1, 64, 2, 3, 65, 0.5),          # batch 1, 65 channels (one channel over the paired-tile limit)
    (7, 64, 7, 7, 65, 0.5),          # twin tiles, odd batch, resident weights
    (1, 512, 7, 7, 70, 0.3),         # batch 1 with 7-pixel rows: no second image to pair with
])
def test_conv_ws_vs_oracle(B, Cin, H, W, Cout, density):
    _case(B, Cin, H, W, Cout, density, seed=B + Cin + H + W + Cout)


def test_conv_ws_epilogue_variants():
    _case(2, 64, 14, 14, 64, 0.3, seed=1, residual=False, relu=True, relu_out=False)
    _case(2, 64, 14, 14, 64, 0.3, seed=2, residual=False, relu=False, relu_out=False, bias_on=False)
    _case(2, 64, 14, 14, 64, 0.3, seed=3, residual=True, relu=True, relu_out=False)


def _signed_sf(rng, n):
    """Factors of both signs, a zero, and large ones that drive most outputs into saturation."""
    sf = rng.uniform(1e-5, 3e-3, n) * rng.choice([-1.0, 1.0], n)
    sf[::7] = 0.0
    sf[3::11] *= 40.0
    return sf


@pytest.mark.parametrize("residual,relu,relu_out", [(False, True, False), (False, False, False), (True, True, True),
                                                     (True, False, False), (False, False, True)])
def test_conv_ws_conversion_free_epilogue(residual, relu, relu_out, monkeypatch):
    """accel_epilogue.acc_bound < 2^22 selects the epilogue without int<->float conversions (ws_epi16<FAST>): same bytes and
    the same saturation count as the oracle, for factors of both signs / zero / saturating, accumulators AT the bound
    (constant +-extreme images against weights of one magnitude), and the general epilogue gives the same answer."""
    from resnet_accel_b200 import _lib
    same = (0.03, 0.03, 0.03)                   # matched scales: the integer residual add (residual mode 4)
    fast = residual                             # only the residual kernels carry the conversion-free variant (api.cu)
    for seed, fill in ((1, None), (2, -128), (3, 127)):
        _case(4, 32, 14, 14, 64, 1.0, seed=seed, residual=residual, relu=relu, relu_out=relu_out, res_scales=same, wmax=100,
              sf_fn=_signed_sf, x_fill=fill, bias_mag=200000, expect_fast=fast)
    _case(2, 64, 56, 56, 64, 0.3, seed=4, residual=residual, relu=relu, relu_out=relu_out, res_scales=same, sf_fn=_signed_sf,
          expect_fast=fast)
    # a bound of 2^22 or more (dense 128-magnitude weights, K = 576) keeps the general epilogue
    _case(2, 64, 14, 14, 64, 1.0, seed=5, residual=residual, relu=relu, relu_out=relu_out, res_scales=same, expect_fast=False)
    monkeypatch.setenv("ACCEL_NO_FAST_EPI", "1")
    _lib.lib().accel_debug_set_timeline(None)
    try:
        _case(4, 32, 14, 14, 64, 1.0, seed=2, residual=residual, relu=relu, relu_out=relu_out, res_scales=same, wmax=100,
              sf_fn=_signed_sf, x_fill=-128, bias_mag=200000, expect_fast=False)
    finally:
        monkeypatch.delenv("ACCEL_NO_FAST_EPI")
        _lib.lib().accel_debug_set_timeline(None)


def test_conv_ws_many_tiles_persistent():
    """More tiles than CTAs: every CTA walks several tiles (accumulator double-buffering, ring wrap-around)."""
    _case(40, 64, 28, 28, 64, 0.3, seed=11)
    _case(24, 256, 14, 14, 128, 0.3, seed=12)


@pytest.mark.parametrize("mode", ["0", "1", "3"])
def test_conv_ws_activation_loaders(mode, monkeypatch):
    """The activation stages arrive by TMA tensor tiles (default, ACCEL_WS_TMA=3) or by the LDGSTS loader warps (0; 1 = TMA for
    64-byte rows only): same bytes either way, for every row width, the stride-2 / fused-downsample and the pointwise modes."""
    from resnet_accel_b200 import _lib
    monkeypatch.setenv("ACCEL_WS_TMA", mode)
    _lib.lib().accel_debug_set_timeline(None)
    try:
        _case(3, 64, 56, 56, 64, 0.3, seed=31)                       # 64-byte rows, paired M = 64 tiles
        _case(3, 128, 28, 28, 128, 0.3, seed=32, residual=False, relu=True, relu_out=False)      # 32-byte rows
        _case(20, 256, 14, 14, 256, 0.3, seed=33)                    # 16-byte rows, streamed weights
        _case(2, 64, 13, 37, 200, 0.4, seed=34)                      # ragged rows and channels
        _case(2, 64, 30, 62, 70, 0.2, seed=35)                       # widest row
        _case(5, 512, 7, 7, 512, 0.3, seed=36)                       # twin tiles: always the LDGSTS loaders
    finally:
        monkeypatch.delenv("ACCEL_WS_TMA")
        _lib.lib().accel_debug_set_timeline(None)


def test_conv_ws_dense_layout_and_fallback():
    """W % 16 == 0 tensors are accepted without padding; unaligned tensors fall back to the gather kernels."""
    import torch
    from resnet_accel_b200 import _lib
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(9)
    B, Cin, H, W, Cout = 2, 32, 10, 30, 20
    Wm = rng.integers(-128, 128, (Cout, Cin * 9), dtype=np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    sf = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32)
    plan = BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    before = _lib.lib().accel_debug_counter(0)
    got = plan.conv(torch.from_numpy(x).cuda(), 3, 1, 1, Cout, out_kind="i8", chan_scale=sf).cpu().numpy()   # rows of 30 bytes
    assert _lib.lib().accel_debug_counter(0) == before
    ref, _ = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, 3, 1, 1, sf=sf)
    assert np.array_equal(got, ref)


def _s2_case(B, Cin, H, W, Cout, density, seed, with_ds=True, expect_ws=True):
    import torch
    from resnet_accel_b200 import _lib, ops
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(seed)

    def make(k):
        K = Cin * k * k
        Wm = rng.integers(-128, 128, (Cout, K), dtype=np.int8)
        nbr, nbc = -(-Cout // 14), -(-K // 14)
        keep = rng.random((nbr, nbc)) < density
        Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :K].astype(np.int8)
        bsr = O.build_bsr_14x14_int8_direct(Wm)
        return bsr, BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])

    bsr3, plan3 = make(3)
    bsr1, plan1 = make(1)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    bias3 = rng.integers(-1000, 1000, Cout, dtype=np.int32)
    bias1 = rng.integers(-300, 300, Cout, dtype=np.int32)
    sf3 = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32)
    sf1 = rng.uniform(1e-3, 8e-3, Cout).astype(np.float32)
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    base_in = xd._base if xd._base is not None else xd
    if base_in.shape[-1] > W:
        base_in[..., W:] = -5
    Ho, Wo = H // 2, W // 2
    out3, out1 = ops.alloc_padded((B, Cout, Ho, Wo)), ops.alloc_padded((B, Cout, Ho, Wo))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    before = _lib.lib().accel_debug_counter(0)
    if with_ds:
        ops.conv_dual(plan3, plan1, xd, Cout, chan_scale=sf3, chan_scale_ds=sf1, bias=bias3, bias_ds=bias1, relu=True,
                      relu_ds=False, out=out3, out_ds=out1, sat_count=cnt)
    else:
        plan3.conv(xd, 3, 2, 1, Cout, "i8", chan_scale=sf3, bias=bias3, relu=True, out=out3, sat_count=cnt)
    torch.cuda.synchronize()
    assert _lib.lib().accel_debug_counter(0) == before + (1 if expect_ws else 0), "unexpected kernel choice"
    ref3, sat3 = c_oracle.conv_bsr_layer(x, bsr3["indptr"], bsr3["indices"], bsr3["data"], Cout, 3, 2, 1, bias=bias3, relu=True, sf=sf3)
    assert np.array_equal(out3.cpu().numpy(), ref3)
    sat = sat3
    if with_ds:
        ref1, sat1 = c_oracle.conv_bsr_layer(x, bsr1["indptr"], bsr1["indices"], bsr1["data"], Cout, 1, 2, 0, bias=bias1, relu=False,
                                             sf=sf1)
        assert np.array_equal(out1.cpu().numpy(), ref1)
        sat += sat1
    assert int(cnt.item()) == sat
    for o in (out3, out1) if with_ds else (out3,):
        base = o._base if o._base is not None else o
        assert int(base[..., Wo:].abs().sum().item()) == 0


@pytest.mark.parametrize("B,Cin,H,W,Cout,density", [
    (2, 64, 56, 56, 128, 0.3),       # layer2.0: conv1 + downsample, one output row per tile
    (3, 128, 28, 28, 256, 0.3),      # layer3.0: two output rows per tile, two channel groups
    (5, 256, 14, 14, 512, 0.3),      # layer4.0: streamed weights, a whole image per tile, one 512-column accumulator set
    (2, 32, 10, 20, 40, 0.6),        # small / odd channel count
    (2, 64, 12, 60, 70, 0.1),        # wide rows, very sparse
    (3, 160, 16, 16, 130, 0.4),      # streamed weights, 8 output rows (the loaders cap the tile at 7), two groups
    (2, 192, 20, 28, 64, 0.3),       # streamed weights, 32-pixel rows (three output rows per tile)
    (1, 160, 6, 40, 48, 0.5),        # streamed weights, 64-pixel rows (one output row per tile)
])
def test_conv_ws_stride2_with_downsample(B, Cin, H, W, Cout, density):
    _s2_case(B, Cin, H, W, Cout, density, seed=B + Cin + H + W)


def test_conv_ws_stride2_alone():
    _s2_case(2, 64, 28, 28, 96, 0.4, seed=5, with_ds=False)
    _s2_case(9, 128, 28, 28, 128, 0.3, seed=6, with_ds=False)
    _s2_case(3, 256, 14, 14, 200, 0.3, seed=7, with_ds=False)     # streamed weights: whole-image tiles, two accumulator sets
    _s2_case(2, 160, 12, 24, 64, 0.5, seed=8, with_ds=False)


def _pw_case(B, Cin, H, W, Cout, density, seed, residual=False, relu=True, relu_out=False, res_scales=(0.05, 0.02, 0.04)):
    """1x1 / stride 1 / pad 0 on the weight-stationary kernel (pointwise mode) vs the C oracle."""
    import torch
    from resnet_accel_b200 import _lib, ops
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(seed)
    Wm = rng.integers(-128, 128, (Cout, Cin)).astype(np.int8)
    nbr, nbc = -(-Cout // 14), -(-Cin // 14)
    keep = rng.random((nbr, nbc)) < density
    keep[:, : min(3, nbc)] = False                   # the first 32-channel chunk holds no block: it still initialises the accumulator
    Wm = Wm * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:Cout, :Cin].astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    bias = rng.integers(-1000, 1000, Cout, dtype=np.int32)
    sf = rng.uniform(1e-4, 2e-3, Cout).astype(np.float32)
    res = rng.integers(-128, 128, (B, Cout, H, W), dtype=np.int8) if residual else None
    plan = BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    xd = ops.alloc_padded(x.shape)
    xd.copy_(torch.from_numpy(x).cuda())
    base_in = xd._base if xd._base is not None else xd
    if base_in.shape[-1] > W:
        base_in[..., W:] = 77                          # garbage in the row padding of the input must not matter
    rd = None
    if residual:
        rd = ops.alloc_padded(res.shape)
        rd.copy_(torch.from_numpy(res).cuda())
    out = ops.alloc_padded((B, Cout, H, W))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    before = _lib.lib().accel_debug_counter(0)
    plan.conv(xd, 1, 1, 0, Cout, out_kind="i8", chan_scale=sf, bias=bias, relu=relu, residual=rd, res_scales=res_scales,
              sat_count=cnt, relu_out=relu_out, out=out)
    torch.cuda.synchronize()
    assert _lib.lib().accel_debug_counter(0) == before + 1, "the weight-stationary kernel did not run"
    ref, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, 1, 1, 0, bias=bias, relu=relu, sf=sf,
                                       residual=res, res_scales=res_scales)
    if relu_out:
        ref = np.maximum(ref, 0)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert int(cnt.item()) == sat
    base = out._base if out._base is not None else out
    assert int(base[..., W:].abs().sum().item()) == 0


@pytest.mark.parametrize("B,Cin,H,W,Cout,density,residual", [
    (2, 64, 56, 56, 64, 0.5, False),          # ResNet-50 layer1.0.conv1: <= 64 channels, paired tiles
    (2, 64, 56, 56, 256, 0.3, True),          # layer1.x.conv3 with the identity added
    (3, 256, 56, 56, 64, 0.3, False),         # layer1.1.conv1
    (3, 512, 28, 28, 128, 0.3, False),        # layer2: 32-pixel rows
    (4, 1024, 14, 14, 256, 0.3, False),       # layer3: weights resident at Cin 1024 (32 chunks)
    (4, 256, 14, 14, 1000, 0.3, True),        # a channel count that is no multiple of 128
    (5, 2048, 7, 7, 512, 0.2, False),         # layer4: two images per staged row, weights streamed (64 chunks), odd batch
    (4, 512, 7, 7, 2048, 0.3, True),          # layer4.x.conv3: 16 channel groups
    (40, 96, 28, 28, 160, 0.4, False),        # many tiles per CTA, 3 chunks
])
def test_conv_ws_pointwise_vs_oracle(B, Cin, H, W, Cout, density, residual):
    _pw_case(B, Cin, H, W, Cout, density, seed=B + Cin + H + Cout, residual=residual, relu=not residual, relu_out=residual)


def test_subsample2_and_strided_pointwise():
    """x[:, :, ::2, ::2] on the GPU, and a 1x1 / stride 2 convolution = the pointwise kernel on the sub-sampled input."""
    import torch
    from resnet_accel_b200 import ops
    from resnet_accel_b200.ops import BsrPlan
    rng = np.random.default_rng(77)
    for shape in ((3, 5, 14, 14), (2, 64, 57, 55), (2, 32, 8, 32)):
        x = torch.from_numpy(rng.integers(-128, 128, shape, dtype=np.int8)).cuda()
        xp = ops.alloc_padded(shape)
        xp.copy_(x)
        for src in (x, xp):
            got = ops.subsample2_int8(src)
            assert torch.equal(got, x[:, :, ::2, ::2])
            base = got._base if got._base is not None else got
            assert int(base[..., got.shape[-1]:].abs().sum().item()) == 0
    B, Cin, H, W, Cout = 2, 256, 28, 28, 512
    Wm = rng.integers(-128, 128, (Cout, Cin)).astype(np.int8)
    bsr = O.build_bsr_14x14_int8_direct(Wm)
    x = rng.integers(-128, 128, (B, Cin, H, W), dtype=np.int8)
    sf = rng.uniform(1e-4, 1e-3, Cout).astype(np.float32)
    plan = BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    xs = ops.subsample2_int8(torch.from_numpy(x).cuda())
    got = plan.conv(xs, 1, 1, 0, Cout, out_kind="i8", chan_scale=sf).cpu().numpy()
    ref, _ = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], Cout, 1, 2, 0, sf=sf)
    assert np.array_equal(got, ref)
