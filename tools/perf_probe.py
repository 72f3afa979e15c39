#!/usr/bin/env python3
"""Developer probe (GPU): per-layer timings of the ResNet-18 BSR network and the 4096^3 sparsity sweep."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from resnet_accel_b200 import exporters as E, layers as L, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    batch = int(os.environ.get("BATCH", 256))
    sparsity = float(os.environ.get("SPARSITY", 70))
    which = os.environ.get("WHICH", "net,gemm")
    if "net" in which:
        t0 = time.time()
        net = L.BsrNetwork(L.resnet18_specs(), sparsity, batch)
        torch.cuda.synchronize()
        print(f"setup {time.time()-t0:.1f}s")
        x = torch.randint(-128, 128, (batch, 3, 224, 224), dtype=torch.int8, device="cuda")
        net.forward(x)
        torch.cuda.synchronize()
        work = net.work()
        tot = 0.0
        xin = ops.alloc_padded(tuple(x.shape))
        xin.copy_(x)
        for sp, wl in zip(net.specs, work["layers"]):
            if sp.name in net.fused_pool.values() or sp.name in net.fused_ds.values():
                continue                      # timed with the layer it is fused into
            prev_name = net.specs[net.specs.index(sp) - 1].name if net.specs.index(sp) else None
            src = xin if sp.name == "conv1" else net.buffers[sp.src or prev_name]

            def run(sp=sp, src=src):
                t = {"input": src}
                t.update(net.buffers)
                Lr = net.layers.get(sp.name)
                if sp.name in net.fused_pool:
                    ops.conv_pool(Lr.plan, src, sp.c_out, chan_scale=Lr.sf, relu=True, out=net.buffers[net.fused_pool[sp.name]])
                    return
                out = net.buffers[sp.name]
                if sp.name in net.fused_ds:
                    D = net.layers[net.fused_ds[sp.name]]
                    ops.conv_dual(Lr.plan, D.plan, src, sp.c_out, chan_scale=Lr.sf, chan_scale_ds=D.sf, relu=True, relu_ds=False,
                                  out=out, out_ds=net.buffers[net.fused_ds[sp.name]])
                    return
                if sp.kind == "conv":
                    if sp.residual:
                        Lr.plan.conv(src, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=Lr.sf, residual=t[sp.residual],
                                     res_scales=(0.05, 0.05, 0.05), out=out, relu_out=True)
                    else:
                        Lr.plan.conv(src, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=Lr.sf, relu=sp.relu, out=out)
                elif sp.kind == "maxpool":
                    ops.maxpool_i8(src, sp.k, sp.stride, sp.pad, out=out)
                elif sp.kind == "avgpool":
                    ops.avgpool_i8(src, out=out)
                else:
                    Lr.plan.gemm(src.reshape(src.shape[0], -1), "i32", n_channels=sp.c_out, out=out)
            ms = timeit(run, 5, 2)
            tot += ms
            tops = wl["ops"] / ms / 1e9
            gbs = wl["bytes"] / ms / 1e6
            mma = net.layers[sp.name].plan.num_mma if sp.name in net.layers else 0
            tag = sp.name + ("+pool" if sp.name in net.fused_pool else "+ds" if sp.name in net.fused_ds else "")
            print(f"{tag:22s} {ms:8.3f} ms  {tops:8.1f} TOPS(useful)  {gbs:8.1f} GB/s  mma/tile={mma}")
        print(f"sum of layers {tot:.3f} ms -> {batch/tot*1e3:.0f} img/s")
        ms = timeit(lambda: net.forward(x), 5, 2)
        print(f"eager forward {ms:.3f} ms -> {batch/ms*1e3:.0f} img/s")
        net.capture(x)
        ms = timeit(lambda: net.replay(), 10, 3)
        print(f"graph replay  {ms:.3f} ms -> {batch/ms*1e3:.0f} img/s ; useful TOPS {work['ops']/ms/1e9:.1f} ; GB/s {work['bytes']/ms/1e6:.1f}")
    if "gemm" in which:
        rng = np.random.default_rng(0)
        W = rng.integers(-128, 128, (4096, 4096), dtype=np.int8)
        A = torch.from_numpy(rng.integers(-128, 128, (4096, 4096), dtype=np.int8)).cuda()
        for pct in (0, 50, 70, 90):
            mask = E.create_sparse_mask((4096, 4096), pct, block_size=14, seed=42)
            bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(W * mask.astype(np.int8)).cuda(), device=True)
            plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
            out = torch.empty((4096, 4102), dtype=torch.int32, device="cuda")
            ms = timeit(lambda: plan.gemm(A, "i32", out=out), 5, 2)
            ops_ = 2 * 4096 * plan.num_blocks * 196
            print(f"gemm4096 sparsity {pct:2d}%: {ms:8.3f} ms  {ops_/ms/1e9:8.1f} useful TOPS  blocks={plan.num_blocks} mma/tile={plan.num_mma}")


if __name__ == "__main__":
    main()
