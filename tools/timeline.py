#!/usr/bin/env python3
"""Developer probe (GPU): per-CTA phase timeline (clock64 stamps) of one ResNet-18 conv layer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

batch = int(os.environ.get("BATCH", 256))
specs = L.resnet18_specs()
for name in os.environ.get("LAYER", "layer1.0.conv1,layer3.1.conv1").split(","):
    idx = [i for i, s in enumerate(s for s in specs if s.kind in ("conv", "fc")) if s.name == name][0]
    sp = [s for s in specs if s.name == name][0]
    lay = L.BsrLayer(sp, L.synthetic_conv_weights(sp, 70.0, idx)["w2"])
    x = ops.alloc_padded((batch, sp.c_in, sp.h, sp.w))
    x.copy_(torch.randint(-128, 128, (batch, sp.c_in, sp.h, sp.w), dtype=torch.int8, device="cuda"))
    y = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    run = lambda: lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, relu=True, out=y)
    for _ in range(2):
        run()
    M = batch * sp.h_out * sp.w_out
    n_cta = (-(-M // 128)) * 8
    buf = torch.zeros(n_cta * 32, dtype=torch.int64, device="cuda")
    _lib.lib().accel_debug_set_timeline(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.lib().accel_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 32)
    t = t[t[:, 0] != 0]
    d = t - t[:, :1]
    names = ["entry", "prologue", "stage0 acquired", "stage0 published", "producers done", "acc complete", "epilogue done", "exit",
             "st2: start", "st2: h_full", "st2: x_empty", "st2: gathered", "st2: st_wait", "st2: published",
             "epi0: top", "epi0: loaded", "-", "epi1: top", "epi1: loaded", "-", "epi2: top", "epi2: loaded", "-", "epi: loop end",
             "ld0: start", "ld0: issued", "ld1: start", "ld1: issued", "ld2: start", "ld2: issued", "ld3: start", "ld3: issued"]
    print(f"{name}: {len(t)} CTAs; median cycles since CTA entry (p10 / p50 / p90):")
    for i, n in enumerate(names):
        if n == "-":
            continue
        col = d[:, i][t[:, i] != 0]
        if not len(col):
            continue
        print(f"   {n:18s} {np.percentile(col,10):9.0f} {np.percentile(col,50):9.0f} {np.percentile(col,90):9.0f}")
