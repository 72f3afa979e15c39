#!/usr/bin/env python3
"""Round-2 additions to tests/golden: fixtures for the generic / ResNet file writers, the hardware blob reader, the random
block masks of the ResNet trainer, relu6, the pruning schedule and the whole-model quantiser - all produced by running the
REFERENCE's own code in the build container (same rules as tools/make_golden.py: nothing here is product code, nothing
copies reference source).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_r2.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("ACCEL_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(REF, "sw"))
sys.path.insert(0, ROOT)


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _read(path):
    with open(path, "rb") as f:
        return np.frombuffer(f.read(), dtype=np.uint8).copy()


def main():
    import torch
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    import training.export_bsr as gen
    rn = _load_by_path("ref_export_resnet18_bsr", os.path.join(REF, "sw/training/export_resnet18_bsr.py"))
    quant = _load_by_path("ref_quantize", os.path.join(REF, "sw", "INT8 quantization", "quantize.py"))
    from oracle import c_oracle
    c_oracle.build(ref=True)
    R = c_oracle.ref()
    import ctypes as C
    rng = np.random.default_rng(2026101802)
    fx = {}
    tmp = tempfile.mkdtemp()

    # ---- 1. generic exporter (export_bsr.py:177-241): block-row-indexed scale quirk, float32 and float64 flavours
    for tag, dt, (rows, cols, b) in (("g32", np.float32, (40, 56, 8)), ("g64", np.float64, (36, 44, 4))):
        w = (rng.standard_normal((rows, cols)) * 0.2).astype(dt)
        nbr, nbc = -(-rows // b), -(-cols // b)
        keep = rng.random((nbr, nbc)) < 0.6
        keep[1] = False                                   # an empty block-row
        w = w * np.repeat(np.repeat(keep, b, 0), b, 1)[:rows, :cols]
        bsr = gen.build_bsr_from_dense(w, b, b)
        scales = np.maximum(np.abs(w).max(axis=1) / 127.0, 1e-12).astype(dt)
        p = os.path.join(tmp, tag + ".bsr")
        gen.save_bsr_binary_int8(bsr, scales, p)
        fx[f"{tag}_w"], fx[f"{tag}_scales"], fx[f"{tag}_block"] = w, scales, np.int32(b)
        fx[f"{tag}_int8_bytes"] = _read(p)
        gen.save_bsr_binary(bsr, p)
        fx[f"{tag}_f32_bytes"] = _read(p)
        pj = os.path.join(tmp, tag + ".json")
        gen.save_bsr_metadata(bsr, pj, layer_name=tag)
        fx[f"{tag}_meta_json"] = _read(pj)
        # 0-d scale
        gen.save_bsr_binary_int8(bsr, np.asarray(0.013, dtype=dt), p)
        fx[f"{tag}_int8_scalar_bytes"] = _read(p)
        # fewer scales than block-rows: scales[0] past the end
        gen.save_bsr_binary_int8(bsr, scales[:2], p)
        fx[f"{tag}_int8_short_bytes"] = _read(p)

    # ---- 2. ResNet exporter (export_resnet18_bsr.py:98-266): headered blob, per-channel scales inside the blocks
    w = (rng.standard_normal((30, 45)) * 0.1).astype(np.float32)
    w[14:28, 0:14] = 0
    bsr = rn.build_bsr_from_dense(w, 14, 14)
    q, sc = rn.quantize_per_channel(w)
    p = os.path.join(tmp, "rn.bsr")
    rn.save_bsr_binary_int8(bsr, sc, p)
    fx["rn_w"], fx["rn_scales"], fx["rn_bytes"], fx["rn_q"] = w, sc, _read(p), q
    rn.save_bsr_binary_int8(bsr, sc[:20], p)              # scales shorter than the channels: those rows stay zero
    fx["rn_short_bytes"] = _read(p)
    pj = os.path.join(tmp, "rn.json")
    rn.save_bsr_metadata(bsr, pj, layer_name="rn")
    fx["rn_meta_json"] = _read(pj)
    fx["rn_layer_config_json"] = np.frombuffer(json.dumps(rn.get_resnet18_layer_config(), sort_keys=True).encode(), np.uint8).copy()

    # ---- 3. C++ hardware blob (bsr_packer.hpp:489-575) through oracle/_ref
    d = rng.integers(-128, 128, (33, 47), dtype=np.int8)
    d[14:28, 14:42] = 0
    out = np.zeros(1 << 16, np.uint8)
    R.ref_serialize_for_hardware.argtypes = [np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS"), C.c_size_t, C.c_size_t,
                                             np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_size_t]
    R.ref_serialize_for_hardware.restype = C.c_long
    n = R.ref_serialize_for_hardware(d, 33, 47, out, out.size)
    assert n > 0, n
    fx["hw_dense"], fx["hw_blob"] = d, out[:n].copy()
    R.ref_deserialize_from_hardware.argtypes = [np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_size_t] + \
        [np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")] * 3 + [np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")]
    R.ref_deserialize_from_hardware.restype = C.c_long
    hdr, rp, ci, dat = np.zeros(3, np.int64), np.zeros(64, np.int64), np.zeros(64, np.int64), np.zeros(64 * 196, np.int8)
    nnz = R.ref_deserialize_from_hardware(fx["hw_blob"], n, hdr, rp, ci, dat)
    assert nnz == hdr[0]
    fx["hw_hdr"], fx["hw_row_ptr"], fx["hw_col_idx"], fx["hw_data"] = hdr, rp[:hdr[1] + 1].copy(), ci[:nnz].copy(), dat[:nnz * 196].copy()
    assert R.ref_deserialize_from_hardware(fx["hw_blob"][:n - 1].copy(), n - 1, hdr.copy(), rp, ci, dat) == -1      # truncated: throws

    # ---- 4. relu6_int8 (golden_models.cpp:323-330)
    x = np.arange(-128, 128, dtype=np.int8)
    for i, s in enumerate((0.05, 0.1, 0.047244, 1.0, 6.0 / 127.0)):
        y = x.copy()
        R.ref_relu6_int8(y, y.size, s)
        fx[f"relu6_{i}_scale"], fx[f"relu6_{i}_out"] = np.float32(s), y
    a = rng.integers(-200000, 200000, 4000, dtype=np.int32)
    b = a.copy()
    R.ref_relu_int32(b, b.size)
    fx["relu32_in"], fx["relu32_out"] = a, b

    # ---- 5. random block masks of the ResNet trainer (train_resnet18.py:69-132); torchvision is present here
    try:
        tr = _load_by_path("ref_train_resnet18", os.path.join(REF, "sw/training/train_resnet18.py"))
        for i, (shape, bs, sp, seed) in enumerate((((30, 50), (14, 14), 0.7, 42), ((16, 8, 3, 3), (4, 4), 0.5, 7),
                                                   ((20, 33), (14, 14), 0.9, 45))):
            m = tr.create_block_sparse_mask(torch.zeros(shape), bs, sp, seed)
            fx[f"bmask_{i}"] = m.numpy()
            fx[f"bmask_{i}_cfg"] = np.array(list(shape) + [0] * (4 - len(shape)) + list(bs) + [int(sp * 100), seed], np.int32)
    except Exception as e:      # pragma: no cover
        print("create_block_sparse_mask fixture skipped:", repr(e))

    # ---- 6. pruning schedule (blocksparse_train.py:141-321): the four pruning phases without the fine-tuning in between
    for m in ("torchvision", "torchvision.datasets", "torchvision.transforms"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    bt = _load_by_path("ref_blocksparse_train", os.path.join(REF, "sw/training/blocksparse_train.py"))
    torch.manual_seed(11)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = torch.nn.Conv2d(3, 12, 3)
            self.fc1 = torch.nn.Linear(64, 40)
            self.fc2 = torch.nn.Linear(40, 16)

    model = Tiny()
    for n in ("conv1", "fc1", "fc2"):
        fx[f"sched_{n}_w"] = getattr(model, n).weight.detach().numpy().copy()
    masks = {n: torch.ones_like(getattr(model, n).weight, dtype=torch.bool) for n in ("conv1", "fc1", "fc2")}
    for phase, sp in enumerate((0.5, 0.7, 0.85, 0.9)):
        npr = bt.prune_blocks_global(model, masks, sp)
        fx[f"sched_{phase}_pruned"] = np.int64(npr)
        for n in ("conv1", "fc1", "fc2"):
            fx[f"sched_{phase}_{n}_mask"] = masks[n].reshape(masks[n].shape[0], -1).numpy().copy()
            fx[f"sched_{phase}_{n}_w"] = getattr(model, n).weight.detach().reshape(masks[n].shape[0], -1).numpy().copy()
    fx["sched_cfg_json"] = np.frombuffer(json.dumps({n: list(bt.layer_block_cfg(n, getattr(model, n))) for n in ("conv1", "fc1", "fc2")}).encode(), np.uint8).copy()

    # ---- 7. whole-model quantiser (quantize.py:162-214)
    torch.manual_seed(3)
    model2 = Tiny()
    qm = quant.quantize_model_per_channel(model2)
    for key, rec in qm.items():
        k = key.replace(".", "_")
        fx[f"qm_{k}_data"] = rec["data"]
        if "scales" in rec:
            fx[f"qm_{k}_scales"] = rec["scales"]
        else:
            fx[f"qm_{k}_scale"] = np.float64(rec["scale"])
        fx[f"qm_{k}_err"] = np.array([rec["error"][e] for e in ("max_error", "mean_error", "mse", "snr_db")], np.float64)
    for n in ("conv1", "fc1", "fc2"):
        fx[f"qm_{n}_weight_fp32"] = getattr(model2, n).weight.detach().numpy().copy()
        fx[f"qm_{n}_bias_fp32"] = getattr(model2, n).bias.detach().numpy().copy()

    np.savez_compressed(os.path.join(OUT, "r2_cases.npz"), **fx)
    mpath = os.path.join(OUT, "MANIFEST.json")
    with open(mpath) as f:
        man = json.load(f)
    man["files"] = sorted(x for x in os.listdir(OUT) if x.endswith(".npz"))
    man["generator_r2"] = "tools/make_golden_r2.py"
    with open(mpath, "w") as f:
        json.dump(man, f, indent=1)
    print("r2_cases.npz", os.path.getsize(os.path.join(OUT, "r2_cases.npz")), "bytes,", len(fx), "arrays")


if __name__ == "__main__":
    main()
