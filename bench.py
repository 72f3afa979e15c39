#!/usr/bin/env python3
"""Headline benchmark: ResNet-18 BSR-INT8 images/sec @70 % block sparsity (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--sparsity S]

One process per GPU (torchrun for N > 1; RANK / LOCAL_RANK / WORLD_SIZE from the environment).
A "step" is one pass of the hot path over one batch of synthetic images: the ResNet-18 conv
backbone + pools + FC, every conv/FC as a 14x14-BSR INT8 layer with fused per-channel requant
(BASELINE.json configs[2]/[3] shape family: batch 256 per GPU, 224x224).  The batch is sharded
by GPU (weak scaling, no data-path collective).  Rank 0 prints ONE JSON line.

  value      images/s, inputs resident in HBM, whole network replayed as one CUDA graph, L2 flushed
             between timed steps, CUDA-event time, max over ranks.
  e2e        same metric through the public API with HOST buffers: pinned-host -> device copy of
             the int8 images and device -> host read of the INT32 logits inside the timed region.
  roofline   the tensor-core convolution / FC launches of one step (conv_ws_kernel for the 3x3 layers, stem_ws_kernel for
             conv1 + max-pool, bsr_tcp_kernel for the FC): algorithmic bytes / event-timed duration vs the measured HBM
             peak (MEASURED_PEAKS.json); `traffic` = DRAM bytes of the same launches from the committed ncu capture
             (profiles/r01_traffic.json, written by tools/ncu_launches.py).
  cpu_baseline  the reference's C++ golden path (oracle/_ref: conv2d_int8_im2col + relu_int32 +
             requantize_int32_to_int8 + add_residual_int8, hw/sim/cpp/src/golden_models.cpp) on a
             bounded sample of the same workload on the host cores.

`--impl reference` times only that CPU path (no GPU code) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ResNet-18 BSR-INT8 images/sec @70% sparsity"
UNIT = "images/s"


# ----------------------------------------------------------------------------------------- helpers
def measured_traffic():
    """DRAM bytes (read + write) of the conv / fc launches of one step, from the committed ncu capture of this command."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), 2.0 * float(d.get("bf16_tflops_sustained", 1400.0)), "measured"
    return 6650.0, 2.0 * 1400.0, "fallback"          # B200_PROFILING.md fallback figures


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nvml_rows, self._stop, self._thr = [], threading.Event(), None

    def _nvml_loop(self):
        """NVML from a thread, every ~2 ms: the timed regions last tens of milliseconds, shorter than nvidia-smi's period."""
        try:
            import pynvml as N
            import torch
            N.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = N.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = N.nvmlDeviceGetHandleByIndex(self.index)
            mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            names = (("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap))
            while not self._stop.is_set():
                sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                bits = int(N.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.nvml_rows.append((sm, mx, [n for n, b in names if bits & b]))
                time.sleep(0.002)
        except Exception:
            pass

    def start(self):
        self._thr = threading.Thread(target=self._nvml_loop, daemon=True)
        self._thr.start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        sm, mx, reasons = [], 0.0, set()
        for v, m, rs in self.nvml_rows:
            sm.append(v); mx = max(mx, m); reasons.update(rs)
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        load = [v for v in sm if v > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU reference arm
def host_network(sparsity: float):
    """Synthetic ResNet-18 weights on the host, built with numpy + the oracle only (no GPU code)."""
    from oracle import bsr_oracle as O
    from resnet_accel_b200.layers import resnet18_specs, S_ACT_IN, S_ACT_OUT
    specs, layers, idx = resnet18_specs(), {}, 0
    for sp in specs:
        if sp.kind not in ("conv", "fc"):
            continue
        seed = 42 + idx
        np.random.seed(seed)
        w4 = np.random.randn(sp.c_out, sp.c_in, sp.k, sp.k).astype(np.float32) * np.sqrt(2.0 / (sp.c_in * sp.k * sp.k))
        w2 = w4.reshape(sp.c_out, -1)
        w2 = (w2 * O.create_sparse_mask(w2.shape, sparsity, 14, seed)).astype(np.float32)
        q, sw = O.quantize_symmetric_per_channel(w2, axis=0)
        layers[sp.name] = {"q": q.reshape(sp.c_out, sp.c_in, sp.k, sp.k), "in_scale": (np.float32(S_ACT_IN) * sw).astype(np.float32)}
        idx += 1
    return specs, layers, S_ACT_OUT


def cpu_forward_image(x, specs, layers, s_out):
    """One image through the network using ONLY reference C++ golden calls (oracle/_ref)."""
    from oracle import c_oracle
    R = c_oracle.ref()
    t = {"input": x}
    prev = "input"
    for sp in specs:
        src = t[sp.src] if sp.src else t[prev]
        if sp.kind == "conv":
            L = layers[sp.name]
            if sp.residual:
                y = c_oracle.ref_conv_layer_image(src, L["q"], None, sp.stride, sp.pad, False, L["in_scale"], s_out,
                                                  t[sp.residual], (s_out, s_out, s_out))
                R.ref_relu_int8(y.reshape(-1), y.size)
            else:
                y = c_oracle.ref_conv_layer_image(src, L["q"], None, sp.stride, sp.pad, sp.relu, L["in_scale"], s_out)
        elif sp.kind == "maxpool":
            xp = np.pad(src, ((0, 0), (sp.pad, sp.pad), (sp.pad, sp.pad)), constant_values=-128)
            y = np.empty((sp.c_out, sp.h_out, sp.w_out), np.int8)
            R.ref_maxpool2d_int8(np.ascontiguousarray(xp), y, xp.shape[1], xp.shape[2], sp.c_out, sp.k, sp.stride)
        elif sp.kind == "avgpool":
            y = np.empty(sp.c_out, np.int8)
            R.ref_avgpool_global_int8(np.ascontiguousarray(src), y, sp.h, sp.w, sp.c_out)
        else:  # fc: logits[n] = sum_k x[k] * W[n, k]  -> matmul_int8(A[1,K], B[K,N])
            L = layers[sp.name]
            Bm = np.ascontiguousarray(L["q"].reshape(sp.c_out, -1).T)
            y = np.empty((1, sp.c_out), np.int32)
            R.ref_matmul_int8(np.ascontiguousarray(src.reshape(1, -1)), Bm, y, 1, Bm.shape[0], sp.c_out)
        t[sp.name] = y
        prev = sp.name
    return t[specs[-1].name]


def cpu_reference_rate(sparsity: float, n_images: int, threads: int, repeats: int = 1):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    if not c_oracle.have_ref():
        raise RuntimeError("oracle/_ref/libref_golden.so missing (built by __graft_entry__.build() in the dev container)")
    specs, layers, s_out = host_network(sparsity)
    rng = np.random.default_rng(0)
    imgs = rng.integers(-128, 128, (n_images, 3, 224, 224), dtype=np.int8)
    times = []
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            list(ex.map(lambda im: cpu_forward_image(im, specs, layers, s_out), imgs))
            times.append(time.perf_counter() - t0)
    return n_images / min(times), times


def run_reference(args, rank: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_images = max(threads, 1) * args.ref_images_per_thread
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_rate(args.sparsity, max(1, threads), threads)
    rate, times = cpu_reference_rate(args.sparsity, n_images, threads, repeats=max(1, args.steps))
    line = {"metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * min(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": "resnet18_full_bsr14_int8", "sparsity_pct": args.sparsity, "image": 224,
                       "images_per_step": n_images},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{n_images} images per step through golden_models.cpp (dense conv2d_int8_im2col path), one image per host thread"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from resnet_accel_b200 import layers as L

    B = args.batch
    net = L.BsrNetwork(L.resnet18_specs(), args.sparsity, B)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x_dev = torch.randint(-128, 128, (B, 3, 224, 224), dtype=torch.int8, device="cuda", generator=gen)
    x_host = x_dev.cpu().pin_memory()
    logits_host = torch.empty((B, 1000), dtype=torch.int32).pin_memory()
    net.capture(x_dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    work = net.work()
    conv_names = [sp.name for sp in net.specs if sp.kind in ("conv", "fc")]
    n_launch_step = net.n_launches

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, K, W):
        for _ in range(W):
            flush.zero_()
            step_fn()
        barrier()
        evs = []
        for _ in range(K):
            flush.zero_()                                   # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (1) kernel-only: inputs resident in HBM, graph replay
    ms_total = timed(lambda: net.replay(), args.steps, args.warmup)
    # (2) end to end through the public API with HOST buffers.  Two-deep pipeline, all inside the timed region:
    #     copy stream: pinned host -> device staging buffer of step i+1   |  compute stream: staging -> network input,
    #     graph replay, INT32 logits -> pinned host.  Events on the compute stream bracket all K steps.
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    staged = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(K):
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        copy_stream.wait_event(e0)                                  # the first copy starts inside the timed region
        for i in range(K):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(consumed[b])             # the staging buffer was read by step i-2
                staging[b].copy_(x_host, non_blocking=True)
                staged[b].record(copy_stream)
            cur.wait_event(staged[b])
            net.static_in.copy_(staging[b], non_blocking=True)
            consumed[b].record(cur)
            net.graph.replay()
            logits_host.copy_(net.buffers["fc"], non_blocking=True)
        e1.record(cur)
        cur.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(2):
        e2e_run(max(args.warmup, 3))
    barrier()
    ms_e2e = e2e_run(args.steps)
    barrier()
    if dist is not None:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    clocks = sampler.stop() if rank == 0 else None

    # (3) dominant kernel: every conv/fc launch of one step, event-timed on the launching stream
    def conv_only():
        net.forward(net.static_in)
    # eager forward = same launches as the graph (the two pools are ~5 % of the step and carry no algorithmic conv bytes)
    for _ in range(2):
        conv_only()
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); conv_only(); e1.record()
        torch.cuda.synchronize()
        kms += e0.elapsed_time(e1)
    kms /= reps
    conv_bytes = sum(l["bytes"] for l in work["layers"] if l["name"] in conv_names)
    conv_ops = sum(l["ops"] for l in work["layers"] if l["name"] in conv_names)
    dense_ops = sum(l["dense_ops"] for l in work["layers"] if l["name"] in conv_names)

    if rank == 0:
        hbm_peak, int8_peak_tops, src = measured_peaks()
        traffic = measured_traffic()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total / 1e3)
        e2e_val = world * B * args.steps / (ms_e2e / 1e3)
        ach_gbs = conv_bytes / (kms / 1e3) / 1e9
        ach_tops = conv_ops / (kms / 1e3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": {"workload": "resnet18_full_bsr14_int8", "sparsity_pct": args.sparsity, "batch_per_gpu": B,
                       "image": 224, "block": 14, "parallelism": f"batch-shard x{world} (no collective)",
                       "l2": "value: flushed between timed steps (256 MiB memset); e2e: per-step working set 1.4 GB >> 126 MB L2",
                       "graph": "one CUDA graph per step", "e2e_pipeline": "2-deep: H2D of step i+1 overlaps step i"},
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": (traffic or {}).get("conv_fc_dram_bytes_per_step"),
                         "traffic_source": (traffic or {}).get("source"), "peak_source": src,
                         "kernel": f"conv_ws_kernel (3x3 layers, downsamples fused) + stem_ws_kernel (conv1 + max-pool) + "
                                   f"bsr_tcp_kernel (fc): {sum(1 for n in conv_names if n not in net.fused_ds.values())} "
                                   "launches per step; algorithmic bytes of the reference's layer sequence (unfused) over "
                                   "the event-timed eager forward",
                         "useful_tops": ach_tops, "tensor_frac_of_2x_bf16_sustained": ach_tops / int8_peak_tops,
                         # the weight-stationary kernels contract every 32-channel chunk that holds a stored block:
                         # the work the tensor pipe does is the dense layer's, of which `useful` is the stored share
                         "dense_equiv_tops": dense_ops / (kms / 1e3) / 1e12,
                         "dense_equiv_tensor_frac": dense_ops / (kms / 1e3) / 1e12 / int8_peak_tops,
                         "kernel_ms_per_step": kms, "algorithmic_bytes_per_step": conv_bytes,
                         "useful_ops_per_step": conv_ops},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel()) * world,
                    "d2h_bytes_per_step": int(logits_host.numel() * 4) * world},
            "gpu_launches": n_launch_step * args.steps,
            "clocks": clocks,
            "sat_count": int(net.sat.item()),
        }
        # CPU baseline (bounded sample) on rank 0 at N=1 only
        if world == 1 and not args.no_cpu_baseline:
            try:
                threads = os.cpu_count() or 1
                n_img = threads * args.ref_images_per_thread
                rate, _ = cpu_reference_rate(args.sparsity, n_img, threads)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                                        "sample": f"{n_img} images through golden_models.cpp (dense conv2d_int8_im2col path), one image per host thread"}
            except Exception as e:  # keep the GPU line even if the prebuilt reference library is absent
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sparsity", type=float, default=70.0)
    ap.add_argument("--ref-images-per-thread", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
