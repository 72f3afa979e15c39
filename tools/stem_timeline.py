#!/usr/bin/env python3
"""Developer probe (GPU): who waits for whom in stem_ws_kernel - clock64 stamps of CTA 0 for conv rows 32..95."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

batch = int(os.environ.get("BATCH", 256))
sp = L.resnet18_specs()[0]
lay = L.BsrLayer(sp, L.synthetic_conv_weights(sp, 70.0, 0)["w2"])
x = ops.alloc_padded((batch, 3, 224, 224))
x.copy_(torch.randint(-128, 128, (batch, 3, 224, 224), dtype=torch.int8, device="cuda"))
out = ops.alloc_padded((batch, 64, 56, 56))
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
run = lambda: ops.conv_pool(lay.plan, x, 64, chan_scale=lay.sf, relu=True, out=out, sat_count=cnt)
for _ in range(3):
    run()
buf = torch.zeros(8 * 64, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
_lib.lib().accel_debug_set_timeline(buf.data_ptr())
run()
torch.cuda.synchronize()
_lib.lib().accel_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(8, 64)
t0 = t[t > 0].min()
names = ["conv: raw ready", "conv: a_full arrive", "issuer: a_full seen", "issuer: row committed", "epi: acc_full seen", "epi: acc_empty arrive",
         "issuer: acc_empty seen", "conv: next row requested"]
print(f"dbg={os.environ.get('ACCEL_DBG_FLAGS', '0')}  cycles since first stamp; row: " + " | ".join(names))
for r in range(0, 40):
    print(f"{32 + r:4d}: " + " ".join(f"{(t[k, r] - t0) if t[k, r] else -1:8d}" for k in range(8)) +
          f"   | convert {t[1, r] - t[0, r]:6d}  mma-issue {t[3, r] - t[2, r]:6d}  epi {t[5, r] - t[4, r]:6d}  acc ready after commit {t[4, r] - t[3, r]:6d}")
