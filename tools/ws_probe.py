#!/usr/bin/env python3
"""Developer probe (GPU): time single ResNet-18 3x3 stride-1 layers on the weight-stationary kernel.
ACCEL_DBG_FLAGS=1 (no epilogue work) / 2 (no MMAs) isolate the pipeline stages."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import _lib, layers as L, ops  # noqa: E402

_lib.lib().accel_debug_set_timeline(None)      # picks up ACCEL_DBG_FLAGS
batch = int(os.environ.get("BATCH", 256))
specs = L.resnet18_specs()
for name in os.environ.get("LAYER", "layer1.0.conv1,layer1.0.conv2,layer2.1.conv1,layer3.1.conv1,layer4.1.conv1").split(","):
    idx = [i for i, s in enumerate(s for s in specs if s.kind in ("conv", "fc")) if s.name == name][0]
    sp = [s for s in specs if s.name == name][0]
    syn = L.synthetic_conv_weights(sp, 70.0, idx)
    lay = L.BsrLayer(sp, syn["w2"])
    x = ops.alloc_padded((batch, sp.c_in, sp.h, sp.w))
    x.copy_(torch.randint(-128, 128, (batch, sp.c_in, sp.h, sp.w), dtype=torch.int8, device="cuda"))
    res = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    res.copy_(torch.randint(-128, 128, (batch, sp.c_out, sp.h_out, sp.w_out), dtype=torch.int8, device="cuda"))
    out = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run():
        if sp.residual:
            lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, residual=res,
                          res_scales=(0.05, 0.05, 0.05), relu_out=True, out=out, sat_count=cnt)
        else:
            lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, relu=True, out=out, sat_count=cnt)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if os.environ.get("GRAPH"):                  # 20 launches replayed from one CUDA graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                run()
        graph.replay()
        torch.cuda.synchronize()
    e0.record()
    if graph is not None:
        graph.replay()
    else:
        for _ in range(20):
            run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:20s} {e0.elapsed_time(e1) / 20 * 1000:8.1f} us  (dbg={os.environ.get('ACCEL_DBG_FLAGS', '0')})")
