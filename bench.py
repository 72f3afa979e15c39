#!/usr/bin/env python3
"""Benchmark of the B200-native BSR-INT8 hot path (BASELINE.json `metric`: ResNet-18 BSR-INT8 images/sec @70 % sparsity).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W] [--batch B] [--sparsity S]

One process per GPU (torchrun for N > 1; RANK / LOCAL_RANK / WORLD_SIZE from the environment).  Rank 0 prints ONE JSON line.

Workloads (`--workload`, default `resnet18` = the configuration the metric is quoted on):
  resnet18             ResNet-18 at 224x224, every conv/FC a 14x14-BSR INT8 layer with fused per-channel requant, batch 256
                       per GPU, batch-sharded over the GPUs (weak scaling, no data-path collective).   BASELINE configs 3 / 4.
  gemm4096             synthetic BSR INT8 GEMM 4096 x 4096 x 4096 at `--sparsity` (SURVEY.md 8d recipe).      BASELINE config 2.
  mnist                MNIST-CNN INT8 with FC1 pruned to `--sparsity` by block L2 norm, batch 64.               BASELINE config 1.
  resnet50_fc_sharded  ResNet-50, batch-sharded trunk, block-row-sharded FC with NCCL all-gathers.              BASELINE config 5.

A "step" is one pass of the hot path over one batch of synthetic input.
  value      units/s, inputs resident in HBM, the step replayed as one CUDA graph (where the workload has one), L2 flushed
             between timed steps, CUDA-event time on the launching stream, max over ranks.
  e2e        the same metric through the public API with HOST buffers: pinned-host -> device copy of the step's input and
             device -> host read of its result inside the timed region (`ResNetInference.run_inference_pipelined`, ...).
  roofline   per-kernel table (`kernels`: name, us, algorithmic bytes, dense-equivalent ops, fractions) from CUDA events
             around every launch of an eager forward, and the totals: algorithmic bytes / time vs the measured HBM peak
             (MEASURED_PEAKS.json), dense-equivalent ops / time vs the measured INT8 tensor peak (profiles/r02_int8_peak.json).
  bit_exact  the step's output for a sample of its inputs compared with the CPU reference, outside the timed region.
  sustained  the same step replayed back to back for `--sustain-seconds` (clocks under a seconds-long load).
  cpu_baseline  the reference's CPU path on a bounded sample of the same workload on the host cores (rank 0, N = 1).

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
WORKLOADS = ("resnet18", "gemm4096", "mnist", "resnet50_fc_sharded")


# ----------------------------------------------------------------------------------------- helpers
def _json(path):
    try:
        with open(os.path.join(ROOT, path)) as f:
            return json.load(f)
    except Exception:
        return None


def measured_peaks():
    """(HBM GB/s, INT8 dense TOPS burst, INT8 dense TOPS sustained, source strings)."""
    d = _json("MEASURED_PEAKS.json")
    hbm, hbm_src = (float(d["hbm_gbs"]), "MEASURED_PEAKS.json (driver, copy bandwidth)") if d and "hbm_gbs" in d else \
        (6650.0, "fallback of B200_PROFILING.md")
    p = _json("profiles/r02_int8_peak.json")
    if p and "int8_tops_burst" in p:
        return hbm, float(p["int8_tops_burst"]), float(p["int8_tops_sustained"]), hbm_src, \
            "profiles/r02_int8_peak.json (cuBLASLt INT8 8192^3 on this pool: best of 10 / 4 s back to back)"
    bf = float(d.get("bf16_tflops", 1590.0)) if d else 1590.0
    bfs = float(d.get("bf16_tflops_sustained", 1400.0)) if d else 1400.0
    return hbm, 2.0 * bf, 2.0 * bfs, hbm_src, "2 x measured bf16 (no INT8 measurement found)"


class ClockSampler:
    """nvidia-smi / NVML clocks and throttle reasons sampled while the timed regions run (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.nvml_rows, self._stop, self._thr = [], threading.Event(), None

    def _nvml_loop(self):
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
                self.nvml_rows.append((time.time(), sm, mx, [n for n, b in names if bits & b]))
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

    def window(self, t0: float, t1: float):
        """Median SM clock of the NVML samples taken inside [t0, t1]."""
        v = [sm for t, sm, _, _ in self.nvml_rows if t0 <= t <= t1]
        return (float(np.median(v)) if v else None), len(v)

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        sm, mx, reasons = [], 0.0, set()
        for _, v, m, rs in self.nvml_rows:
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


def bind_to_gpu_numa(local_rank: int):
    """Pin this process to the CPUs NVML names as local to its GPU (its NUMA node), BEFORE any pinned host buffer is
    allocated: first-touch then places those buffers on the same node.  Returns a short description for the JSON line."""
    try:
        import pynvml as N
        import torch
        N.nvmlInit()
        try:
            h = N.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)).encode())
        except Exception:
            h = N.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = N.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"cpus": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})", "bound": True}
        return {"cpus": None, "bound": False}
    except Exception as e:      # no NVML / not permitted: run unbound
        return {"cpus": None, "bound": False, "why": repr(e)[:80]}


def h2d_ceiling_gbs(nbytes: int, seconds: float = 0.5) -> float:
    """Bare pinned-host -> device copy loop of this rank (all ranks run it at the same time): the host-side ceiling that
    bounds `e2e`, measured with nothing else on the GPU."""
    import torch
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(8):
            dst.copy_(src, non_blocking=True)
        n += 8
        torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    return n * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9


# ----------------------------------------------------------------------------------------- CPU reference arm (ResNet-18)
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


def cpu_reference_rate(sparsity: float, n_images: int, threads: int, repeats: int = 1, images=None):
    """(images/s, per-repeat seconds, logits of the last repeat) of the reference C++ golden chain, one image per thread."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    if not c_oracle.have_ref():
        raise RuntimeError("oracle/_ref/libref_golden.so missing (built by __graft_entry__.build() in the dev container)")
    specs, layers, s_out = host_network(sparsity)
    imgs = images if images is not None else np.random.default_rng(0).integers(-128, 128, (n_images, 3, 224, 224), dtype=np.int8)
    times, outs = [], None
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            outs = list(ex.map(lambda im: cpu_forward_image(im, specs, layers, s_out), imgs))
            times.append(time.perf_counter() - t0)
    return len(imgs) / min(times), times, np.stack([o.reshape(-1) for o in outs])


# ----------------------------------------------------------------------------------------- workload descriptions
def workload_config(args, world: int) -> dict:
    """The `config` object - identical for the GPU arm and the reference arm."""
    w = args.workload
    if w == "resnet18":
        return {"workload": "resnet18_full_bsr14_int8", "sparsity_pct": args.sparsity, "batch_per_gpu": args.batch, "image": 224,
                "block": 14, "parallelism": f"batch-shard x{world} (no collective)"}
    if w == "gemm4096":
        return {"workload": "bsr_gemm_4096x4096x4096_int8", "sparsity_pct": args.sparsity, "rows_per_gpu": 4096, "block": 14,
                "parallelism": f"row-shard x{world} (no collective)"}
    if w == "mnist":
        return {"workload": "mnist_cnn_int8_fc1_bsr14", "fc1_sparsity_pct": args.sparsity, "batch_per_gpu": 64, "block": 14,
                "parallelism": f"batch-shard x{world} (no collective)"}
    return {"workload": "resnet50_bsr14_int8_fc_block_row_sharded", "sparsity_pct": args.sparsity, "batch_per_gpu": args.batch,
            "image": 224, "block": 14,
            "parallelism": f"trunk batch-shard x{world}; fc block-rows x{world} + 2 NCCL all-gathers (features, logits)"}


def workload_metric(args):
    w = args.workload
    if w == "resnet18":
        return (METRIC if abs(args.sparsity - 70.0) < 1e-9 else f"ResNet-18 BSR-INT8 images/sec @{args.sparsity:g}% sparsity"), UNIT
    if w == "gemm4096":
        return f"BSR-INT8 GEMM 4096x4096x4096 @{args.sparsity:g}% block sparsity, GEMMs/sec", "GEMMs/s"
    if w == "mnist":
        return f"MNIST-CNN INT8 (FC1 {args.sparsity:g}% block-sparse) images/sec", UNIT
    return f"ResNet-50 BSR-INT8 images/sec @{args.sparsity:g}% sparsity, block-row-sharded FC", UNIT


# ----------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    metric, unit = workload_metric(args)
    threads = os.cpu_count() or 1
    cfg = workload_config(args, world)
    if args.workload == "resnet18":
        n_images = max(threads, 1) * args.ref_images_per_thread
        for _ in range(args.warmup if args.warmup < 2 else 1):
            cpu_reference_rate(args.sparsity, max(1, threads), threads)
        rate, times, _ = cpu_reference_rate(args.sparsity, n_images, threads, repeats=max(1, args.steps))
        kind, sample = "reference", (f"{n_images} images per step through golden_models.cpp (dense conv2d_int8_im2col path), "
                                     "one image per host thread")
    else:
        rate, times, kind, sample = cpu_port_rate(args, threads)
    line = {"metric": metric, "value": rate, "unit": unit, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * min(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_port_rate(args, threads: int):
    """CPU arm of the secondary workloads: the plain-C / NumPy port of the reference algorithm (the reference's own Python
    golden for them is a pure-Python triple loop, ~1 us per MAC) on a bounded sample."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import bsr_oracle as O
    from oracle import c_oracle
    if args.workload == "gemm4096":
        rng = np.random.default_rng(0)
        W = rng.integers(-128, 128, (4096, 4096), dtype=np.int8)
        W = (W * O.create_sparse_mask((4096, 4096), args.sparsity, 14, 42).astype(np.int8)).astype(np.int8)
        A = rng.integers(-128, 128, (4096, 4096), dtype=np.int8)
        bsr = O.build_bsr_14x14_int8_direct(W)
        rows_per_thread = 8
        rows = threads * rows_per_thread
        chunks = [A[i * rows_per_thread:(i + 1) * rows_per_thread] for i in range(threads)]
        times = []
        with ThreadPoolExecutor(max_workers=threads) as ex:
            for _ in range(max(1, min(args.steps, 3))):
                t0 = time.perf_counter()
                list(ex.map(lambda a: c_oracle.bsr_gemm_i32(a, bsr["indptr"], bsr["indices"], bsr["data"]), chunks))
                times.append(time.perf_counter() - t0)
        return (rows / 4096.0) / min(times), times, "port", (f"{rows} of 4096 activation rows through oracle/c (gemm_bsr_int8_golden "
                                                            "restated in C), scaled linearly, one row slice per host thread")
    if args.workload == "mnist":
        w = dict(np.load(os.path.join(ROOT, "tests", "golden", "mnist_int8.npz")))
        imgs = np.concatenate([w["inputs_u8"], np.random.default_rng(64).integers(0, 256, (32, 28, 28), dtype=np.uint8)], 0)
        sc = O.mnist_activation_scales(O.mnist_float_forward(O.mnist_preprocess(imgs), w))
        fc1 = O.mnist_prune_fc1(w["fc1_weight_int8"], args.sparsity / 100.0)
        times = []
        for _ in range(max(1, min(args.steps, 5))):
            t0 = time.perf_counter()
            O.mnist_cnn_int8_forward(imgs, w, sc, fc1)
            times.append(time.perf_counter() - t0)
        return 64 / min(times), times, "port", "the 64-image batch through the NumPy restatement of the golden chain, one thread"
    raise SystemExit("--impl reference: no CPU arm for workload " + args.workload)


# ----------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    # before the first pinned allocation; a single rank keeps every host core (the CPU baseline runs on them)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"cpus": None, "bound": False, "why": "single rank"}
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W = build_workload(args, rank, local_rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, K, Wn):
        for _ in range(Wn):
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
        return allmax(sum(a.elapsed_time(b) for a, b in evs))

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # (1) kernel-only: inputs resident in HBM
    ms_total = timed(W.step, args.steps, args.warmup)
    # (2) end to end through the public API with HOST buffers, all copies inside the timed region
    for _ in range(2):
        W.e2e(max(args.warmup, 3))
    barrier()
    ms_e2e = allmax(W.e2e(args.steps))
    barrier()
    # (3) sustained: the step back to back for seconds (clocks under load)
    sustained = None
    if args.sustain_seconds > 0:
        n_est = max(8, int(args.sustain_seconds * 1e3 / max(ms_total / args.steps, 1e-3)))
        barrier()
        t_w0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_est):
            W.step()
        e1.record(); torch.cuda.synchronize()
        t_w1 = time.time()
        ms_s = allmax(e0.elapsed_time(e1))
        mhz, ns = sampler.window(t_w0, t_w1) if rank == 0 else (None, 0)
        sustained = {"seconds": ms_s / 1e3, "steps": n_est, "value": world * W.units_per_step * n_est / (ms_s / 1e3), "unit": W.unit,
                     "sm_mhz_median": mhz, "clock_samples": ns, "l2": "not flushed (back to back)"}
    # (4) host-copy ceiling of the box, all ranks at once
    ceiling = None
    if W.h2d_bytes:
        barrier()
        g = h2d_ceiling_gbs(W.h2d_bytes)
        if dist is not None:
            t = torch.tensor([g], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            g = float(t.item())
        ceiling = g
    clocks = sampler.stop() if rank == 0 else None
    # (5) per-kernel table + bit-exactness, outside every timed region
    kernels = W.kernel_table(flush)
    bit_exact, cpu_base = W.check_and_cpu_baseline(rank, world)

    if rank == 0:
        hbm_peak, i8_burst, i8_sus, hbm_src, i8_src = measured_peaks()
        ms_step = ms_total / args.steps
        value = world * W.units_per_step * args.steps / (ms_total / 1e3)
        e2e_val = world * W.units_per_step * args.steps / (ms_e2e / 1e3)
        k_us = sum(k["us"] for k in kernels)
        k_bytes = sum(k["bytes"] for k in kernels)
        k_ops = sum(k["useful_ops"] for k in kernels)
        k_dense = sum(k["dense_ops"] for k in kernels)
        for k in kernels:
            k["hbm_frac"] = k["bytes"] / (k["us"] * 1e-6) / 1e9 / hbm_peak if k["us"] else None
            k["dense_equiv_tensor_frac"] = k["dense_ops"] / (k["us"] * 1e-6) / 1e12 / i8_burst if k["us"] else None
        ach_gbs = k_bytes / (k_us * 1e-6) / 1e9 if k_us else 0.0
        dense_tops = k_dense / (k_us * 1e-6) / 1e12 if k_us else 0.0
        hbm_frac, tensor_frac = ach_gbs / hbm_peak, dense_tops / i8_burst
        bound = "tensor" if tensor_frac >= hbm_frac else "hbm"
        metric, unit = workload_metric(args)
        cfg = workload_config(args, world)
        cfg.update({"l2": "value: flushed between timed steps (256 MiB memset); e2e: per-step working set >> 126 MB L2"
                    if args.workload in ("resnet18", "resnet50_fc_sharded") else
                    "value: flushed between timed steps (256 MiB memset)", **W.config_extra})
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": cfg,
            "roofline": {"bound": bound,
                         "achieved": dense_tops if bound == "tensor" else ach_gbs,
                         "peak": i8_burst if bound == "tensor" else hbm_peak,
                         "unit": "TOP/s" if bound == "tensor" else "GB/s",
                         "frac": max(hbm_frac, tensor_frac),
                         "hbm": {"achieved_gbs": ach_gbs, "peak_gbs": hbm_peak, "frac": hbm_frac, "peak_source": hbm_src},
                         "tensor": {"dense_equiv_tops": dense_tops, "useful_tops": k_ops / (k_us * 1e-6) / 1e12 if k_us else 0.0,
                                    "peak_tops_burst": i8_burst, "peak_tops_sustained": i8_sus, "dense_equiv_frac": tensor_frac,
                                    "useful_frac": (k_ops / (k_us * 1e-6) / 1e12 / i8_burst) if k_us else 0.0, "peak_source": i8_src},
                         "traffic": W.measured_traffic(), "traffic_source": W.traffic_source,
                         "kernel": W.kernel_note, "kernel_ms_per_step": k_us / 1e3, "algorithmic_bytes_per_step": k_bytes,
                         "useful_ops_per_step": k_ops, "dense_equiv_ops_per_step": k_dense, "kernels": kernels},
            "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": int(W.h2d_bytes) * world,
                    "d2h_bytes_per_step": int(W.d2h_bytes) * world, "api": W.e2e_api,
                    "h2d_ceiling_gbs_all_ranks": ceiling, "h2d_needed_gbs_all_ranks": W.h2d_bytes * world / (ms_step / 1e3) / 1e9,
                    "numa": numa},
            "gpu_launches": W.launches_per_step * args.steps,
            "bit_exact": bit_exact,
            "clocks": clocks,
            "sustained": sustained,
        }
        line.update(W.extra_line())
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------- workloads (GPU arm)
class _Workload:
    unit = UNIT
    config_extra: dict = {}
    traffic_source = None
    kernel_note = ""
    e2e_api = ""
    h2d_bytes = 0
    d2h_bytes = 0

    def measured_traffic(self):
        return None

    def extra_line(self):
        return {}


def _time_eager(fn, flush, reps=3):
    """Per-launch (name, us) from CUDA events recorded on the launching stream around every launch of an eager forward."""
    import torch
    acc = {}
    order = []
    for r in range(reps + 1):
        flush.zero_()
        ev = []
        fn(ev)
        torch.cuda.synchronize()
        if r == 0:
            continue                # warm-up
        for name, a, b in ev:
            if name not in acc:
                acc[name] = 0.0
                order.append(name)
            acc[name] += a.elapsed_time(b) * 1e3
    return [(n, acc[n] / reps) for n in order]


class ResNet18Workload(_Workload):
    def __init__(self, args, rank, world):
        import torch
        from resnet_accel_b200 import layers as L
        self.args, self.world = args, world
        B = args.batch
        self.engine = L.ResNetInference(batch=B)
        self.engine.load_synthetic(args.sparsity)
        self.net = self.engine.net
        gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
        self.x_dev = torch.randint(-128, 128, (B, 3, 224, 224), dtype=torch.int8, device="cuda", generator=gen)
        self.x_host = [self.x_dev.cpu().pin_memory(), self.x_dev.cpu().pin_memory()]      # one pinned buffer per in-flight step
        self.logits_host = [torch.empty((B, 1000), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.net.capture(self.x_dev)
        self.units_per_step = B
        self.h2d_bytes, self.d2h_bytes = int(self.x_dev.numel()), B * 1000 * 4
        self.launches_per_step = self.net.n_launches
        self.config_extra = {"graph": "one CUDA graph per step", "e2e_pipeline": "2-deep: H2D of step i+1 overlaps step i"}
        self.e2e_api = "ResNetInference.run_inference_pipelined(pinned host int8 batches -> pinned host int32 logits)"
        self.kernel_note = ("conv_ws_kernel (3x3 layers, stride-2 layers with their downsample fused) + stem_ws_kernel (conv1 + "
                            "max-pool) + avgpool + gemm_ws_kernel (fc); algorithmic bytes of the reference's layer sequence "
                            "(unfused), dense-equivalent ops = the dense layer's 2*MACs")
        t = _json("profiles/r02_traffic.json") or _json("profiles/r01_traffic.json")
        self._traffic = t
        self.traffic_source = (t or {}).get("source")

    def step(self):
        self.net.replay()

    def e2e(self, K):
        import torch
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        self.engine.run_inference_pipelined([self.x_host[i & 1] for i in range(K)], [self.logits_host[i & 1] for i in range(K)])
        e1.record(cur)
        cur.synchronize()
        return e0.elapsed_time(e1)

    def measured_traffic(self):
        return (self._traffic or {}).get("conv_fc_dram_bytes_per_step")

    def kernel_table(self, flush):
        work = {l["name"]: l for l in self.net.work()["layers"]}
        rows = _time_eager(lambda ev: self.net.forward(self.net.static_in, launch_events=ev), flush)
        fused = dict(self.net.fused_ds)
        fused.update(self.net.fused_pool)
        out = []
        for name, us in rows:
            parts = [name] + ([fused[name]] if name in fused else [])
            out.append({"name": "+".join(parts), "us": us, "bytes": sum(work[p]["bytes"] for p in parts),
                        "useful_ops": sum(work[p]["ops"] for p in parts), "dense_ops": sum(work[p]["dense_ops"] for p in parts)})
        return out

    def check_and_cpu_baseline(self, rank, world):
        """GPU logits of a few of the step's own images against the reference C++ golden chain; the same CPU run is the
        cpu_baseline sample at N = 1."""
        if rank != 0:
            return None, None
        threads = len(os.sched_getaffinity(0)) or 1
        n_img = min(threads * self.args.ref_images_per_thread if world == 1 else min(4, threads), self.args.batch)
        if self.args.no_cpu_baseline:
            return None, None
        try:
            imgs = self.x_host[0][:n_img].numpy()
            rate, _, ref_logits = cpu_reference_rate(self.args.sparsity, n_img, threads, images=imgs)
            self.net.replay(self.x_dev)
            got = self.net.buffers["fc"][:n_img].cpu().numpy()
            ok = bool(np.array_equal(got, ref_logits))
            base = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                    "sample": f"{n_img} images through golden_models.cpp (dense conv2d_int8_im2col path), one image per host thread"}
            return {"ok": ok, "checked": f"INT32 logits of {n_img} images of the timed batch vs oracle/_ref (reference C++ golden)"}, \
                (base if world == 1 else None)
        except Exception as e:      # keep the GPU line even if the prebuilt reference library is absent
            return {"ok": None, "checked": f"unavailable: {e}"}, None

    def extra_line(self):
        return {"sat_count": int(self.net.sat.item())}


class Gemm4096Workload(_Workload):
    unit = "GEMMs/s"

    def __init__(self, args, rank, world):
        import torch
        from resnet_accel_b200 import exporters as E, ops
        n = 4096
        rng = np.random.default_rng(0)
        Wm = rng.integers(-128, 128, (n, n), dtype=np.int8)
        Wm = (Wm * E.create_sparse_mask((n, n), args.sparsity, block_size=14, seed=42).astype(np.int8)).astype(np.int8)
        self.A = rng.integers(-128, 128, (n, n), dtype=np.int8)
        self.bsr = E.build_bsr_14x14_int8_direct(torch.from_numpy(Wm).cuda(), device=True)
        self.plan = ops.BsrPlan(self.bsr["indptr"], self.bsr["indices"], self.bsr["data"], n_block_cols=self.bsr["num_block_cols"])
        self.x = torch.from_numpy(self.A).cuda()
        self.out = torch.empty((n, self.plan.n_out_padded), dtype=torch.int32, device="cuda")
        self.x_host = torch.from_numpy(self.A).pin_memory()
        self.out_host = torch.empty((n, self.plan.n_out_padded), dtype=torch.int32).pin_memory()
        self.x_stage = torch.empty_like(self.x)
        self.units_per_step, self.launches_per_step = 1, 1
        self.h2d_bytes, self.d2h_bytes = n * n, n * self.plan.n_out_padded * 4
        nb = self.plan.num_blocks
        self.work = {"bytes": n * n + nb * 196 + 4 * (self.plan.n_block_rows + 1) + 4 * nb + n * self.plan.n_out_padded * 4,
                     "ops": 2 * n * nb * 196, "dense": 2 * n * n * n}
        self.e2e_api = "BsrPlan.gemm on a staged copy of the pinned host activations; INT32 result copied back to pinned host"
        self.kernel_note = "gemm_ws_kernel<2, 0>: dense-equivalent CTA-pair GEMM, INT32 output"
        self.config_extra = {"output": "int32 [4096, 4102]"}
        self.args = args

    def step(self):
        self.plan.gemm(self.x, "i32", out=self.out)

    def e2e(self, K):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            self.x_stage.copy_(self.x_host, non_blocking=True)
            self.plan.gemm(self.x_stage, "i32", out=self.out)
            self.out_host.copy_(self.out, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def kernel_table(self, flush):
        import torch
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); self.step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return [{"name": "gemm_ws_kernel", "us": float(np.median(ts)), "bytes": self.work["bytes"], "useful_ops": self.work["ops"],
                 "dense_ops": self.work["dense"]}]

    def check_and_cpu_baseline(self, rank, world):
        if rank != 0 or self.args.no_cpu_baseline:
            return None, None
        from oracle import c_oracle
        import torch
        rows = np.sort(np.random.default_rng(1).choice(4096, 64, replace=False))
        rp, ci, blk = (self.bsr[k].cpu().numpy() for k in ("indptr", "indices", "data"))
        t0 = time.perf_counter()
        ref = c_oracle.bsr_gemm_i32(self.A[rows], rp, ci, blk)
        dt = time.perf_counter() - t0
        self.step()
        ok = bool(np.array_equal(self.out[torch.from_numpy(rows).cuda()].cpu().numpy(), ref))
        base = {"value": (64 / 4096.0) / dt, "unit": self.unit, "cores": 1, "kind": "port",
                "sample": "64 of 4096 activation rows through oracle/c (gemm_bsr_int8_golden restated in C), scaled linearly"}
        return {"ok": ok, "checked": "64 sampled output rows vs oracle/c"}, (base if world == 1 else None)


class MnistWorkload(_Workload):
    def __init__(self, args, rank, world):
        import torch
        from resnet_accel_b200.mnist import MnistCnnInt8
        self.args = args
        self.w = dict(np.load(os.path.join(ROOT, "tests", "golden", "mnist_int8.npz")))
        self.imgs = np.concatenate([self.w["inputs_u8"], np.random.default_rng(64).integers(0, 256, (32, 28, 28), dtype=np.uint8)], 0)
        self.net = MnistCnnInt8(self.w, batch=64, fc1_sparsity=args.sparsity / 100.0)
        self.scales = self.net.calibrate(self.imgs)
        self.net.build()
        self.net.run(self.imgs)
        self.host_in = torch.from_numpy(self.imgs).pin_memory()
        self.host_out = torch.empty((64, 10), dtype=torch.float32).pin_memory()
        self.units_per_step, self.launches_per_step = 64, 6
        self.h2d_bytes, self.d2h_bytes = 64 * 28 * 28, 64 * 10 * 4
        self.e2e_api = "MnistCnnInt8.run(pinned host uint8 images) -> logits copied to pinned host"
        self.kernel_note = "bsr_tc / bsr_tcp kernels (conv1, conv2, fc1, fc2) + maxpool; launch-latency-bound at batch 64"
        self.config_extra = {"graph": "one CUDA graph per step"}

    def step(self):
        self.net.graph.replay()

    def e2e(self, K):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            y = self.net.run(self.host_in.cuda(non_blocking=True))
            self.host_out.copy_(y, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def kernel_table(self, flush):
        import torch
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); self.step(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        wk = self.net.work()
        return [{"name": "mnist_cnn_graph (6 launches)", "us": float(np.median(ts)), "bytes": wk["bytes"], "useful_ops": wk["ops"],
                 "dense_ops": wk["ops"]}]

    def check_and_cpu_baseline(self, rank, world):
        if rank != 0 or self.args.no_cpu_baseline:
            return None, None
        from oracle import bsr_oracle as O
        fc1 = O.mnist_prune_fc1(self.w["fc1_weight_int8"], self.args.sparsity / 100.0) if self.args.sparsity > 0 else \
            O.build_bsr_14x14_int8_direct(self.w["fc1_weight_int8"])
        t0 = time.perf_counter()
        ref = O.mnist_cnn_int8_forward(self.imgs, self.w, self.scales, fc1)
        dt = time.perf_counter() - t0
        got = self.net.run(self.imgs).cpu().numpy()
        ok = bool(np.array_equal(got, ref["logits"]))
        base = {"value": 64 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "the 64-image batch through the NumPy restatement of the golden chain"}
        return {"ok": ok, "checked": "de-quantised logits of the 64-image batch vs the oracle chain"}, (base if world == 1 else None)


class ResNet50ShardedWorkload(_Workload):
    def __init__(self, args, rank, world):
        import torch
        from resnet_accel_b200 import layers as L
        self.args, self.rank, self.world = args, rank, world
        B = args.batch
        self.net = L.ShardedFcNetwork(L.resnet50_specs(), args.sparsity, B)
        gen = torch.Generator(device="cuda").manual_seed(77 + rank)
        self.x = L.ops.alloc_padded((B, 3, 224, 224))
        self.x.copy_(torch.randint(-128, 128, (B, 3, 224, 224), dtype=torch.int8, device="cuda", generator=gen))
        self.x_host = self.x.cpu().pin_memory()
        self.logits_host = torch.empty((world * B, 1000), dtype=torch.int32).pin_memory()
        self.net.forward(self.x)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()            # the trunk (no collective inside) is one graph; the head stays eager
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.net.trunk_forward(self.x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.net.trunk_forward(self.x)
        self.units_per_step = B
        self.launches_per_step = self.net.trunk.n_launches + 1
        self.h2d_bytes, self.d2h_bytes = int(self.x.numel()), world * B * 1000 * 4
        self.e2e_api = "pinned host images -> trunk graph -> feature all-gather -> sharded FC -> logits all-gather -> pinned host"
        self.kernel_note = "ResNet-50 trunk (conv_ws 3x3 + gather kernels for the 1x1 layers) + gemm_ws FC shard; see collective_us"
        self.config_extra = {"graph": "trunk = one CUDA graph; head (2 all-gathers + FC shard) eager on the same stream"}
        self._coll = None

    def step(self):
        self.graph.replay()
        self.net.head_forward()

    def e2e(self, K):
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            self.x.copy_(self.x_host, non_blocking=True)
            self.graph.replay()
            y = self.net.head_forward()
            self.logits_host.copy_(y, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def kernel_table(self, flush):
        import torch
        from resnet_accel_b200 import parallel
        work = {l["name"]: l for l in self.net.trunk.work()["layers"]}
        rows = _time_eager(lambda ev: self.net.trunk.forward(self.x, launch_events=ev), flush)
        fused = dict(self.net.trunk.fused_ds); fused.update(self.net.trunk.fused_pool)
        out = []
        for name, us in rows:
            parts = [name] + ([fused[name]] if name in fused else [])
            out.append({"name": "+".join(parts), "us": us, "bytes": sum(work[p]["bytes"] for p in parts),
                        "useful_ops": sum(work[p]["ops"] for p in parts), "dense_ops": sum(work[p]["dense_ops"] for p in parts)})
        # the head, piece by piece: feature all-gather, FC shard, logits all-gather
        def t(fn, n=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e3 / n
        net = self.net
        mine = net.features[net.rank * net.batch:(net.rank + 1) * net.batch]
        us_feat = t(lambda: parallel.gather_rows(mine, net.features, net.group))
        us_fc = t(lambda: net.fc.local_gemm(net.features, "i32", bias=net.fc_bias))
        buf = net.fc.local_gemm(net.features, "i32", bias=net.fc_bias)
        us_log = t(lambda: net.fc.gather(buf))
        M, K = net.features.shape
        nb = net.fc.plan.num_blocks if net.fc.plan is not None else 0
        out.append({"name": "fc (block-row shard)", "us": us_fc, "bytes": M * K + nb * 196 + M * net.fc.wmax * 4,
                    "useful_ops": 2 * M * nb * 196, "dense_ops": 2 * M * K * net.fc.wmax})
        self._coll = {"feature_all_gather_us": us_feat, "feature_bytes_total": int(M * K), "logits_all_gather_us": us_log,
                      "logits_bytes_total": int(net.fc.world * net.fc.wmax * M * 4), "fc_shard_us": us_fc,
                      "logits_result": "view of the gather buffer (no copy)" if net.fc.even else "one column copy per shard"}
        return out

    def check_and_cpu_baseline(self, rank, world):
        """Bit-exactness of the sharded head: the FC of ALL feature rows on one GPU (unsharded plan) vs the gathered logits."""
        import torch
        from resnet_accel_b200 import ops
        b = self.net.fc_bsr
        full = ops.BsrPlan(b["indptr"], b["indices"], b["data"], n_block_cols=b["num_block_cols"])
        self.step()
        want = full.gemm(self.net.features, "i32", n_channels=1000, bias=self.net.fc_bias)
        got = self.net.fc.result(self.net._buf)
        ok = bool(torch.equal(got, want))
        if world > 1:
            import torch.distributed as dist
            t_ = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MIN)
            ok = bool(t_.item())
        if rank != 0:
            return None, None
        return {"ok": ok, "checked": "gathered logits of every rank vs the unsharded FC of the gathered features (single plan)"}, None

    def extra_line(self):
        return {"collective": self._coll}


def build_workload(args, rank, local_rank, world):
    return {"resnet18": ResNet18Workload, "gemm4096": Gemm4096Workload, "mnist": MnistWorkload,
            "resnet50_fc_sharded": ResNet50ShardedWorkload}[args.workload](args, rank, world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="resnet18", choices=WORKLOADS)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--sparsity", type=float, default=None)
    ap.add_argument("--ref-images-per-thread", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {"resnet18": 256, "resnet50_fc_sharded": 128}.get(args.workload, 64)
    if args.sparsity is None:
        args.sparsity = 90.0 if args.workload == "mnist" else 70.0
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
