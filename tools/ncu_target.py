#!/usr/bin/env python3
"""Small fixed workload for ncu: a few launches of the tcgen05 kernel.
WHICH=gemm (4096^3 @ SPARSITY %) and/or conv (LAYER of ResNet-18 at BATCH, 70 % sparse)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import exporters as E, layers as L, ops  # noqa: E402

which = os.environ.get("WHICH", "gemm,conv")
reps = int(os.environ.get("REPS", 3))
if "gemm" in which:
    rng = np.random.default_rng(0)
    W = rng.integers(-128, 128, (4096, 4096), dtype=np.int8)
    A = torch.from_numpy(rng.integers(-128, 128, (4096, 4096), dtype=np.int8)).cuda()
    mask = E.create_sparse_mask((4096, 4096), float(os.environ.get("SPARSITY", 50)), block_size=14, seed=42)
    bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(W * mask.astype(np.int8)).cuda(), device=True)
    plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
    out = torch.empty((4096, 4102), dtype=torch.int32, device="cuda")
    for _ in range(reps):
        plan.gemm(A, "i32", out=out)
    torch.cuda.synchronize()
if "conv" in which:
    batch = int(os.environ.get("BATCH", 64))
    specs = L.resnet18_specs()
    for name in os.environ.get("LAYER", "layer1.0.conv1,layer3.1.conv1").split(","):
        idx = [i for i, s in enumerate(s for s in specs if s.kind in ("conv", "fc")) if s.name == name][0]
        sp = [s for s in specs if s.name == name][0]
        syn = L.synthetic_conv_weights(sp, 70.0, idx)
        lay = L.BsrLayer(sp, syn["w2"])
        x = ops.alloc_padded((batch, sp.c_in, sp.h, sp.w))            # 16-byte aligned rows: the TMA / persistent path
        x.copy_(torch.randint(-128, 128, (batch, sp.c_in, sp.h, sp.w), dtype=torch.int8, device="cuda"))
        res = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
        res.copy_(torch.randint(-128, 128, (batch, sp.c_out, sp.h_out, sp.w_out), dtype=torch.int8, device="cuda"))
        out = ops.alloc_padded((batch, sp.c_out, sp.h_out, sp.w_out))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")        # the network runner counts saturations: same variant
        for _ in range(reps):
            if sp.residual:
                y = lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, residual=res,
                                  res_scales=(0.05, 0.05, 0.05), relu_out=True, out=out, sat_count=cnt)
            else:
                y = lay.plan.conv(x, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, relu=True, out=out, sat_count=cnt)
        torch.cuda.synchronize()
print("ok")
