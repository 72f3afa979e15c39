#!/usr/bin/env python3
"""Developer probe (GPU): time the fused stride-2 convolution + downsample calls of ResNet-18 (layerN.0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

_lib.lib().accel_debug_set_timeline(None)
batch = int(os.environ.get("BATCH", 256))
specs = L.resnet18_specs()
names = [s.name for s in specs if s.kind in ("conv", "fc")]
for st in (2, 3, 4):
    sp = [s for s in specs if s.name == f"layer{st}.0.conv1"][0]
    dp = [s for s in specs if s.name == f"layer{st}.0.downsample"][0]
    lay = L.BsrLayer(sp, L.synthetic_conv_weights(sp, 70.0, names.index(sp.name))["w2"])
    dlay = L.BsrLayer(dp, L.synthetic_conv_weights(dp, 70.0, names.index(dp.name))["w2"])
    x = ops.alloc_padded((batch, sp.c_in, sp.h, sp.w))
    x.copy_(torch.randint(-128, 128, (batch, sp.c_in, sp.h, sp.w), dtype=torch.int8, device="cuda"))
    o1 = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    o2 = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run():
        ops.conv_dual(lay.plan, dlay.plan, x, sp.c_out, chan_scale=lay.sf, chan_scale_ds=dlay.sf, relu=True, relu_ds=False,
                      out=o1, out_ds=o2, sat_count=cnt)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"layer{st}.0 conv1+downsample {e0.elapsed_time(e1) / 20 * 1000:8.1f} us  (dbg={os.environ.get('ACCEL_DBG_FLAGS', '0')})")
    if os.environ.get("ALONE"):
        def run2():
            lay.plan.conv(x, 3, 2, 1, sp.c_out, "i8", chan_scale=lay.sf, relu=True, out=o1, sat_count=cnt)
        for _ in range(3):
            run2()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            run2()
        e1.record()
        torch.cuda.synchronize()
        print(f"layer{st}.0 conv1 alone (stride 2) {e0.elapsed_time(e1) / 20 * 1000:8.1f} us")
