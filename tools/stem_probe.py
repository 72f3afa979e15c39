#!/usr/bin/env python3
"""Developer probe (GPU): time the fused stem (conv1 + max-pool) of ResNet-18 at batch 256."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

_lib.lib().accel_debug_set_timeline(None)      # picks up ACCEL_DBG_FLAGS

batch = int(os.environ.get("BATCH", 256))
specs = L.resnet18_specs()
sp = specs[0]
lay = L.BsrLayer(sp, L.synthetic_conv_weights(sp, 70.0, 0)["w2"])
x = ops.alloc_padded((batch, 3, 224, 224))
x.copy_(torch.randint(-128, 128, (batch, 3, 224, 224), dtype=torch.int8, device="cuda"))
out = ops.alloc_padded((batch, 64, 56, 56))
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for use_cnt in (True, False):
    def run():
        ops.conv_pool(lay.plan, x, 64, chan_scale=lay.sf, relu=True, out=out, sat_count=cnt if use_cnt else None)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"stem conv1+maxpool {e0.elapsed_time(e1) / 20 * 1000:8.1f} us  (sat_count={'on' if use_cnt else 'off'}, dbg={os.environ.get('ACCEL_DBG_FLAGS', '0')})")
