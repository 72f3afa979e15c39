#!/usr/bin/env python3
"""Parity + timing of the BSR GEMM on BASELINE config 2 (synthetic 4096^3 sweep, SURVEY.md 8d recipe) and a few FC shapes.

    python tools/gemm_probe.py [--tag T] [--sizes 4096] [--pcts 0,50,70,90] [--iters 20] [--no-check]

Environment switches of the library are read at load time (ACCEL_NO_GEMM_WS, ACCEL_GEMM_WS_CG, ACCEL_GEMM_WS_STAGES), so one
process measures one variant.  Prints one JSON line per case.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def sweep_case(n, pct):
    """W[out, in] int8, block mask create_sparse_mask(seed 42), A int8 - the recipe of SURVEY.md 8d."""
    from resnet_accel_b200 import exporters as E
    rng = np.random.default_rng(0)
    W = rng.integers(-128, 128, (n, n), dtype=np.int8)
    mask = E.create_sparse_mask((n, n), pct, block_size=14, seed=42)
    W = (W * mask.astype(np.int8)).astype(np.int8)
    A = rng.integers(-128, 128, (n, n), dtype=np.int8)
    return W, A


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="")
    ap.add_argument("--sizes", default="4096")
    ap.add_argument("--pcts", default="0,50,70,90")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    import torch
    from resnet_accel_b200 import _lib, exporters as E, ops
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for n in [int(s) for s in args.sizes.split(",")]:
        for pct in [float(s) for s in args.pcts.split(",")]:
            W, A = sweep_case(n, pct)
            bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(W).cuda(), device=True)
            plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
            x = torch.from_numpy(A).cuda()
            out = torch.empty((n, plan.n_out_padded), dtype=torch.int32, device="cuda")
            g0 = L.accel_debug_counter(1)
            plan.gemm(x, "i32", out=out)
            torch.cuda.synchronize()
            used_ws = L.accel_debug_counter(1) - g0
            rec = {"tag": args.tag, "n": n, "sparsity_pct": pct, "blocks": plan.num_blocks, "gemm_ws": int(used_ws),
                   "live_chunks_cg2": int(L.accel_plan_gemm_ws_live_chunks(plan._h, 2)),
                   "env": {k: v for k, v in os.environ.items() if k.startswith("ACCEL_")}}
            if not args.no_check:
                rows = np.random.default_rng(1).choice(n, 256, replace=False)
                ref = (A[rows].astype(np.float64) @ W.T.astype(np.float64)).astype(np.int64)      # exact: |sum| < 2^53
                got = out[torch.from_numpy(rows).cuda()].cpu().numpy()
                rec["rows_ok"] = bool(np.array_equal(got[:, :n], ref) and not got[:, n:].any())
                gen = ops.bsr_gemm_generic(x, bsr["indptr"], bsr["indices"], bsr["data"], plan.n_out_padded)
                rec["full_ok_vs_generic"] = bool(torch.equal(gen, out))
                rec["checksum"] = int(out.to(torch.int64).sum().item())
            # timing: CUDA events on the launching stream, L2 flushed between iterations
            for _ in range(3):
                plan.gemm(x, "i32", out=out)
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.gemm(x, "i32", out=out); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                plan.gemm(x, "i32", out=out)
            e1.record(); torch.cuda.synchronize()
            useful = 2.0 * n * plan.num_blocks * 196
            bytes_alg = n * n + plan.num_blocks * 196 + 4 * (plan.n_block_rows + 1) + 4 * plan.num_blocks + n * plan.n_out_padded * 4
            us = float(np.median(ts))
            rec.update({"us_flushed_median": us, "us_flushed_min": float(min(ts)), "us_back_to_back": e0.elapsed_time(e1) * 1e3 / args.iters,
                        "useful_tops": useful / us / 1e6, "dense_equiv_tops": 2.0 * n * n * n / us / 1e6,
                        "alg_gbs": bytes_alg / us / 1e3})
            print(json.dumps(rec), flush=True)
            del plan, out
    # FC-like shapes (ResNet-18 fc, ResNet-50 fc at batch 1024, MNIST fc1 at batch 64)
    for (M, N, K, pct) in ((256, 1000, 512, 70.0), (1024, 1000, 2048, 70.0), (64, 128, 9216, 90.0)):
        rng = np.random.default_rng(7)
        W = rng.integers(-128, 128, (N, K), dtype=np.int8)
        W = (W * E.create_sparse_mask((N, K), pct, block_size=14, seed=42).astype(np.int8)).astype(np.int8)
        A = rng.integers(-128, 128, (M, K), dtype=np.int8)
        bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(W).cuda(), device=True)
        plan = ops.BsrPlan(bsr["indptr"], bsr["indices"], bsr["data"], n_block_cols=bsr["num_block_cols"])
        x = torch.from_numpy(A).cuda()
        g0 = L.accel_debug_counter(1)
        out = plan.gemm(x, "i32")
        used_ws = L.accel_debug_counter(1) - g0
        ok = None
        if not args.no_check:
            ref = (A.astype(np.float64) @ W.T.astype(np.float64)).astype(np.int64)
            ok = bool(np.array_equal(out.cpu().numpy()[:, :N], ref))
        for _ in range(3):
            plan.gemm(x, "i32", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            plan.gemm(x, "i32", out=out)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"tag": args.tag, "fc": [M, N, K], "sparsity_pct": pct, "gemm_ws": int(used_ws), "ok": ok,
                          "us_back_to_back": e0.elapsed_time(e1) * 1e3 / 50}), flush=True)


if __name__ == "__main__":
    main()
