#!/usr/bin/env python3
"""Developer probe (GPU): per-CTA phase timeline of conv_ws_kernel for ResNet-18 layers (clock64 stamps, 16 per CTA)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

batch = int(os.environ.get("BATCH", 256))
specs = L.resnet18_specs()
NAMES = ["entry", "prologue done", "epilogue past griddep", "epilogue done", "issuer: weights", "issuer: first stage", "issuer: item 0 issued",
         "issuer done", "loaders past griddep", "loaders done", "all roles done", "exit"]
for name in os.environ.get("LAYER", "layer1.0.conv1,layer2.1.conv1,layer3.1.conv1,layer4.1.conv1").split(","):
    idx = [i for i, s in enumerate(s for s in specs if s.kind in ("conv", "fc")) if s.name == name][0]
    sp = [s for s in specs if s.name == name][0]
    lay = L.BsrLayer(sp, L.synthetic_conv_weights(sp, 70.0, idx)["w2"])
    x = ops.alloc_padded((batch, sp.c_in, sp.h, sp.w))
    x.copy_(torch.randint(-128, 128, (batch, sp.c_in, sp.h, sp.w), dtype=torch.int8, device="cuda"))
    y = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    run = lambda: lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, relu=True, out=y)
    for _ in range(3):
        run()
    buf = torch.zeros(148 * 128, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    _lib.lib().accel_debug_set_timeline(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    _lib.lib().accel_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 128)
    t = t[t[:, 0] != 0]
    d = (t[:, :12] - t[:, :1]).astype(np.float64)
    wall_ns = t[:, 13].max() - t[:, 14].min()
    print(f"{name}: {len(t)} CTAs, event time {e0.elapsed_time(e1)*1e3:.1f} us, first entry -> last exit {wall_ns/1e3:.1f} us, "
          f"items/CTA {t[:,12].mean():.1f}, dbg={os.environ.get('ACCEL_DBG_FLAGS','0')}; cycles since CTA entry (p10 / p50 / p90):")
    for i, n in enumerate(NAMES):
        col = d[:, i][t[:, i] != 0]
        if len(col):
            print(f"   {n:26s} {np.percentile(col,10):9.0f} {np.percentile(col,50):9.0f} {np.percentile(col,90):9.0f}")
    ent = (t[:, 14] - t[:, 14].min()) / 1e3
    print(f"   CTA entry spread (us): p50 {np.percentile(ent,50):.1f}  p90 {np.percentile(ent,90):.1f}  max {ent.max():.1f}")

    # per-stage dynamics of stages 32..63 of CTA 0: when the loader got the slot, when the issuer saw the data
    c0 = t[0]
    L_, I_, J_ = c0[16:48] - c0[0], c0[64:96] - c0[0], c0[96:128] - c0[0]
    print("   stage: loader got slot / loader issued / issuer saw data   | issue - slot, data - issue  (cycles since entry, CTA 0)")
    for i in range(0, 32, int(os.environ.get("TL_STEP", 2))):
        print(f"   {32+i:4d}: {L_[i]:8d} {J_[i]:8d} {I_[i]:8d}   | {J_[i]-L_[i]:6d} {I_[i]-J_[i]:6d}")
