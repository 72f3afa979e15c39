#!/usr/bin/env python3
"""Measured dense INT8 tensor peak of this B200, same protocol as MEASURED_PEAKS.json uses for bf16:
cuBLASLt INT8 GEMM (torch._int_mm, int8 x int8 -> int32) 8192^3, best of 10 (burst) and back to back for 4 s (sustained),
clocks sampled by NVML during the sustained loop.  Writes one JSON object (stdout, and --out)."""
import argparse
import json
import sys
import threading
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    n = args.n
    a = torch.randint(-128, 128, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 128, (n, n), dtype=torch.int8, device="cuda").t().contiguous().t()   # column-major B, as cuBLASLt wants
    ops = 2.0 * n ** 3
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch._int_mm(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    clocks, stop = [], threading.Event()

    def sample():
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
            while not stop.is_set():
                clocks.append((N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), N.nvmlDeviceGetPowerUsage(h) / 1e3,
                               int(N.nvmlDeviceGetCurrentClocksEventReasons(h))))
                time.sleep(0.05)
        except Exception as e:      # noqa
            clocks.append(("nvml unavailable", str(e)))
    th = threading.Thread(target=sample, daemon=True); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, iters = time.time(), 0
    e0.record()
    while time.time() - t0 < args.seconds:
        for _ in range(20):
            torch._int_mm(a, b)
        iters += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join(timeout=1)
    sus_ms = e0.elapsed_time(e1) / iters
    sm = [c[0] for c in clocks if isinstance(c[0], (int, float))]
    rec = {"what": f"torch._int_mm (cuBLASLt INT8, s8 x s8 -> s32) {n}^3", "int8_tops_burst": ops / best / 1e9,
           "int8_tops_sustained": ops / sus_ms / 1e9, "burst_ms": best, "sustained_ms": sus_ms, "sustained_iters": iters,
           "sm_mhz_median_sustained": (sorted(sm)[len(sm) // 2] if sm else None),
           "power_w_max": (max(c[1] for c in clocks if isinstance(c[0], (int, float))) if sm else None),
           "gpu": torch.cuda.get_device_name(0)}
    s = json.dumps(rec)
    print(s, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
