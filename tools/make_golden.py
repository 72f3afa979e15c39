#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REFERENCE's own code in the build container.

Run once (needs /root/reference, which exists only in the build container; the GPU box
only ever sees the committed fixtures):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

Sources of truth used:
  * reference Python, imported from /root/reference/sw (golden GEMMs, packers, masks,
    quantisers, host BSRMatrix, pruner with a stubbed matplotlib import);
  * reference C++ golden (hw/sim/cpp/src/golden_models.cpp + bsr_packer.hpp) through
    oracle/_ref/libref_golden.so (built by oracle/Makefile from where the sources lie);
  * reference data files (data/int8, data/bsr_export_14x14, data/fixtures, sw/golden/*.npy).

Nothing here is product code; nothing here copies reference source.
"""
import hashlib
import importlib.util
import json
import os
import sys
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


def main():
    os.makedirs(OUT, exist_ok=True)
    from golden.gemm_bsr_int8 import gemm_bsr_int8
    import golden.golden_fc1_test as gfc1
    from training.export_bsr import build_bsr_from_dense
    from training.export_bsr_14x14 import build_bsr_14x14, build_bsr_14x14_int8_direct
    from exporters.export_transformer import create_sparse_mask
    from host.memory import BSRMatrix, pack_activations
    quant = _load_by_path("ref_quantize", os.path.join(REF, "sw", "INT8 quantization", "quantize.py"))
    from oracle import c_oracle

    rng = np.random.default_rng(20261018)
    manifest = {}

    # ---------------------------------------------------------------- 1. real MNIST INT8 model
    mn = {}
    for l in ("conv1", "conv2", "fc1", "fc2"):
        mn[f"{l}_weight_int8"] = np.load(os.path.join(REF, "data/int8", f"{l}_weight_int8.npy"))
        mn[f"{l}_weight_scales"] = np.load(os.path.join(REF, "data/int8", f"{l}_weight_scales.npy"))
        mn[f"{l}_bias_int8"] = np.load(os.path.join(REF, "data/int8", f"{l}_bias_int8.npy"))
        with open(os.path.join(REF, "data/int8", f"{l}_bias_scale.json")) as f:
            js = json.load(f)
        mn[f"{l}_bias_scale"] = np.float64(js["scale"] if isinstance(js, dict) else js)
    mn["inputs_u8"] = np.load(os.path.join(REF, "sw/golden/mnist_inputs.npy"))
    mn["logits_fp32"] = np.load(os.path.join(REF, "sw/golden/mnist_logits_fp32.npy"))
    mn["tiles_A"] = np.load(os.path.join(REF, "data/int8/tiles/A.npy"))
    mn["tiles_B"] = np.load(os.path.join(REF, "data/int8/tiles/B.npy"))
    with open(os.path.join(REF, "data/int8/tiles/scales.json")) as f:
        mn["tiles_scales_json"] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "mnist_int8.npz"), **mn)

    # shipped 14x14 export: structure + digests (the packer must reproduce them byte for byte)
    exp = {}
    for l in ("conv1", "conv2", "fc1", "fc2"):
        d = os.path.join(REF, "data/bsr_export_14x14", l)
        exp[f"{l}_row_ptr"] = np.load(os.path.join(d, "row_ptr.npy"))
        exp[f"{l}_col_idx"] = np.load(os.path.join(d, "col_idx.npy"))
        raw = open(os.path.join(d, "weights.bsr"), "rb").read()
        exp[f"{l}_sha256"] = np.frombuffer(hashlib.sha256(raw).hexdigest().encode(), dtype=np.uint8)
        with open(os.path.join(d, "weights.meta.json")) as f:
            exp[f"{l}_meta_json"] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "bsr_export_14x14.npz"), **exp)

    # ---------------------------------------------------------------- 2. FC1 known answer
    fc1 = gfc1.load_bsr_layer(os.path.join(REF, "data/bsr_export_14x14/fc1"))
    K = fc1["shape"][1]
    act = ((np.arange(K) % 256) - 128).astype(np.int8).reshape(1, K)
    C1 = gfc1.gemm_bsr_int8_golden(act, fc1)
    np.savez_compressed(os.path.join(OUT, "fc1_known_answer.npz"), activations=act, output=C1)
    manifest["fc1_known_answer_sha256_16"] = hashlib.sha256(C1.tobytes()).hexdigest()[:16]

    # ---------------------------------------------------------------- 3. random INT32 golden cases
    cases = {}
    specs = [  # (M, N_out, K_in, density)
        (3, 28, 42, 0.6), (5, 30, 45, 0.5), (2, 14, 14, 1.0), (4, 56, 70, 0.0), (1, 100, 33, 0.3), (7, 29, 57, 0.8),
    ]
    for i, (M, N, K, dens) in enumerate(specs):
        W = rng.integers(-128, 128, (N, K), dtype=np.int8)
        nbr, nbc = -(-N // 14), -(-K // 14)
        keep = rng.random((nbr, nbc)) < dens
        Wm = W * np.repeat(np.repeat(keep, 14, 0), 14, 1)[:N, :K].astype(np.int8)
        bsr = build_bsr_14x14_int8_direct(Wm)
        layer = {"row_ptr": bsr["indptr"], "col_idx": bsr["indices"], "weights": bsr["data"], "block_h": 14,
                 "block_w": 14}
        for kvar, Kact in (("padK", bsr["padded_shape"][1]), ("rawK", K)):
            A = rng.integers(-128, 128, (M, Kact), dtype=np.int8)
            Cg = gfc1.gemm_bsr_int8_golden(A, layer)
            cases[f"c{i}_{kvar}_A"] = A
            cases[f"c{i}_{kvar}_C"] = Cg
        cases[f"c{i}_W"] = Wm
        cases[f"c{i}_row_ptr"] = bsr["indptr"]
        cases[f"c{i}_col_idx"] = bsr["indices"]
        cases[f"c{i}_blocks"] = bsr["data"]
    # extreme-value cases from hw/sim/cpp/tests/test_stress.cpp:87-245 (intents in comments there)
    for tag, aval, wval in (("max", 127, 127), ("min", -128, -128), ("mixed", 127, -128)):
        W = np.full((28, 28), wval, dtype=np.int8)
        A = np.full((2, 28), aval, dtype=np.int8)
        bsr = build_bsr_14x14_int8_direct(W)
        layer = {"row_ptr": bsr["indptr"], "col_idx": bsr["indices"], "weights": bsr["data"], "block_h": 14,
                 "block_w": 14}
        cases[f"x_{tag}_A"], cases[f"x_{tag}_W"] = A, W
        cases[f"x_{tag}_C"] = gfc1.gemm_bsr_int8_golden(A, layer)
    np.savez_compressed(os.path.join(OUT, "golden_i32_cases.npz"), **cases)

    # ---------------------------------------------------------------- 4. FP32 compat golden (gemm_bsr_int8)
    fp = {}
    for i, (M, K, N, b, f32scales) in enumerate([(4, 64, 8, 8, True), (4, 28, 42, 14, True), (3, 32, 16, 8, False)]):
        Bt = (rng.standard_normal((K, N)) * 3).astype(np.float32)
        keep = rng.random((-(-K // b), -(-N // b))) < 0.7
        Bt = Bt * np.repeat(np.repeat(keep, b, 0), b, 1)[:K, :N].astype(np.float32)
        bsr = build_bsr_from_dense(Bt, b, b)
        A = rng.integers(-128, 128, (M, K), dtype=np.int8)
        scales = (np.abs(rng.standard_normal(max(K, N))) * 0.05 + 0.01)
        scales = scales.astype(np.float32) if f32scales else scales.astype(np.float64)
        sA = np.float32(0.03) if f32scales else 0.03
        Cf = gemm_bsr_int8(A, bsr, sA, scales)
        fp[f"f{i}_A"], fp[f"f{i}_B"], fp[f"f{i}_scales"], fp[f"f{i}_scaleA"] = A, Bt, scales, np.asarray(sA)
        fp[f"f{i}_block"], fp[f"f{i}_C"] = np.int32(b), Cf
        fp[f"f{i}_indptr"], fp[f"f{i}_indices"], fp[f"f{i}_data"] = bsr["indptr"], bsr["indices"], bsr["data"]
    np.savez_compressed(os.path.join(OUT, "golden_fp32_cases.npz"), **fp)

    # ---------------------------------------------------------------- 5. packers / masks / quantisers
    pk = {}
    Wf = (rng.standard_normal((40, 75)) * 0.1).astype(np.float32)
    mask = create_sparse_mask(Wf.shape, 60.0, block_size=14, seed=7)
    Wf = Wf * mask
    pk["mask_40x75_60_s7"] = mask
    pk["mask_64x64_50_b8_s42"] = create_sparse_mask((64, 64), 50.0, block_size=8, seed=42)
    pk["mask_4096_70_b14_s42_blocks"] = np.packbits(
        create_sparse_mask((4096, 4096), 70.0, block_size=14, seed=42)[::14, ::14].astype(np.uint8))
    q, sc = quant.quantize_symmetric_per_channel(Wf, axis=0)
    pk["W_f32"], pk["W_q"], pk["W_scales"] = Wf, q, sc
    qt, st = quant.quantize_symmetric_per_tensor(Wf)
    pk["W_qt"], pk["W_st"] = qt, np.float64(st)
    for name, bsr in (("f32", build_bsr_14x14(Wf)), ("q", build_bsr_14x14(Wf, quantize=True, scale=sc)),
                      ("direct", build_bsr_14x14_int8_direct(q)), ("generic8", build_bsr_from_dense(Wf, 8, 8)),
                      ("generic4x8", build_bsr_from_dense(Wf, 4, 8))):
        pk[f"{name}_data"], pk[f"{name}_indices"], pk[f"{name}_indptr"] = bsr["data"], bsr["indices"], bsr["indptr"]
        pk[f"{name}_meta"] = np.array([bsr["padded_shape"][0], bsr["padded_shape"][1], bsr["num_blocks"],
                                       bsr["num_block_rows"], bsr["num_block_cols"]], dtype=np.int64)
    hb = BSRMatrix.from_dense(q, block_size=14)
    pk["host_row_ptr"], pk["host_col_idx"], pk["host_values"] = hb.row_ptr, hb.col_idx, hb.values
    pk["host_dma"] = np.frombuffer(hb.pack_for_dma(), dtype=np.uint8)
    pk["host_pack_act"] = np.frombuffer(pack_activations(q[:5, :30]), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "packer_cases.npz"), **pk)

    # data/fixtures (generic 8x8): structure + int8 payload of one layer, with its scales
    fx = {}
    d = os.path.join(REF, "data/fixtures/mlp/fc_512_128")
    with open(os.path.join(d, "weights.meta.json")) as f:
        meta = json.load(f)
    fx["row_ptr"], fx["col_idx"] = np.array(meta["row_ptr"], np.int32), np.array(meta["col_idx"], np.int32)
    fx["blocksize"] = np.array(meta["blocksize"], np.int32)
    fx["shape"] = np.array(meta["shape"], np.int32)
    fx["weights_int8"] = np.fromfile(os.path.join(d, "weights_int8.bsr"), dtype=np.int8)
    fx["scales"] = np.load(os.path.join(d, "scales.npy"))
    np.savez_compressed(os.path.join(OUT, "fixture_mlp_512_128.npz"), **fx)

    # ---------------------------------------------------------------- 6. pruner (blocksparse_train.py)
    try:
        for m in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.datasets", "torchvision.transforms"):
            if m not in sys.modules:
                try:
                    __import__(m)
                except Exception:
                    sys.modules[m] = types.ModuleType(m)
        import torch
        bt = _load_by_path("ref_blocksparse_train", os.path.join(REF, "sw/training/blocksparse_train.py"))
        torch.manual_seed(5)

        class Tiny(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.conv1 = torch.nn.Conv2d(2, 8, 3)
                self.fc1 = torch.nn.Linear(40, 24)

        model = Tiny()
        pr = {"conv1_w": model.conv1.weight.detach().numpy().copy(), "fc1_w": model.fc1.weight.detach().numpy().copy()}
        n1, _, _ = bt.compute_block_norms(model.conv1.weight.data, 4, 4)
        n2, _, _ = bt.compute_block_norms(model.fc1.weight.data, 8, 8)
        pr["conv1_norms"], pr["fc1_norms"] = n1.numpy(), n2.numpy()
        masks = {"conv1": torch.ones_like(model.conv1.weight, dtype=torch.bool),
                 "fc1": torch.ones_like(model.fc1.weight, dtype=torch.bool)}
        npr = bt.prune_blocks_global(model, masks, 0.6)
        pr["n_pruned"] = np.int64(npr)
        pr["conv1_mask"] = masks["conv1"].reshape(8, -1).numpy()
        pr["fc1_mask"] = masks["fc1"].numpy()
        np.savez_compressed(os.path.join(OUT, "pruner_cases.npz"), **pr)
    except Exception as e:  # pragma: no cover
        print("pruner fixture skipped:", repr(e))

    # ---------------------------------------------------------------- 7. C++ golden (oracle/_ref)
    c_oracle.build(ref=True)
    R = c_oracle.ref()
    cc = {}
    acc = np.concatenate([np.array([1000, -500, 0, 2000, 5, 15, -5, -15, 25, 35], np.int32),
                          rng.integers(-200000, 200000, 4000, dtype=np.int32)])
    for i, (si, so) in enumerate([(1.0, 10.0), (0.02 * 0.0031, 0.05), (0.0173, 0.41), (3e-5, 7e-3)]):
        out = np.empty(acc.size, np.int8)
        R.ref_requantize_int32_to_int8(acc, out, acc.size, si, so)
        cc[f"rq{i}_scales"], cc[f"rq{i}_out"] = np.array([si, so], np.float32), out
    cc["rq_acc"] = acc
    a = rng.integers(-128, 128, 5000, dtype=np.int8)
    b = rng.integers(-128, 128, 5000, dtype=np.int8)
    for i, (s1, s2, s3) in enumerate([(1.0, 1.0, 1.0), (0.05, 0.02, 0.04), (0.013, 0.027, 0.019)]):
        out = np.empty(a.size, np.int8)
        R.ref_add_residual_int8(a, b, out, a.size, s1, s2, s3)
        cc[f"res{i}_scales"], cc[f"res{i}_out"] = np.array([s1, s2, s3], np.float32), out
    cc["res_a"], cc["res_b"] = a, b
    x = rng.integers(-128, 128, (6, 12, 10), dtype=np.int8)
    for pool, stride in ((2, 2), (3, 2), (3, 1)):
        Ho, Wo = (12 - pool) // stride + 1, (10 - pool) // stride + 1
        out = np.empty((6, Ho, Wo), np.int8)
        R.ref_maxpool2d_int8(x, out, 12, 10, 6, pool, stride)
        cc[f"maxpool_{pool}_{stride}"] = out
    cc["pool_x"] = x
    xa = rng.integers(-128, 128, (9, 7, 7), dtype=np.int8)
    xa[0] = -128; xa[1] = 127; xa[2] = -1
    out = np.empty(9, np.int8)
    R.ref_avgpool_global_int8(xa, out, 7, 7, 9)
    cc["avg_x"], cc["avg_out"] = xa, out
    r8 = np.array([-128, -1, 0, 1, 50, 100, 127, -50], np.int8)
    R.ref_relu_int8(r8, r8.size)
    cc["relu8_out"] = r8
    for i, (Cin, H, W, Cout, k, s, p) in enumerate([(3, 6, 6, 4, 3, 1, 1), (2, 9, 7, 5, 3, 2, 1), (4, 8, 8, 6, 1, 2, 0),
                                                    (1, 10, 10, 3, 3, 1, 0), (3, 11, 11, 2, 7, 2, 3)]):
        xi = rng.integers(-128, 128, (Cin, H, W), dtype=np.int8)
        w = rng.integers(-128, 128, (Cout, Cin, k, k), dtype=np.int8)
        bias = rng.integers(-1024, 1024, Cout, dtype=np.int32)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        o1 = np.empty((Cout, Ho, Wo), np.int32)
        o2 = np.empty((Cout, Ho, Wo), np.int32)
        R.ref_conv2d_int8_im2col(xi, w, bias.ctypes.data, o1, Cin, H, W, Cout, k, s, p)
        R.ref_conv2d_int8_simple(xi, w, bias.ctypes.data, o2, Cin, H, W, Cout, k, s, p)
        assert np.array_equal(o1, o2)
        col = np.empty((Cin * k * k, Ho * Wo), np.int8)
        R.ref_im2col_int8(xi, col, Cin, H, W, k, s, p, Ho, Wo)
        cc[f"conv{i}_x"], cc[f"conv{i}_w"], cc[f"conv{i}_bias"], cc[f"conv{i}_out"] = xi, w, bias, o1
        cc[f"conv{i}_geom"] = np.array([k, s, p], np.int32)
        cc[f"conv{i}_col"] = col
    # Convention A: pack_to_bsr + bsr_matmul_int8 (one zeroed block, non-multiple-of-14 dims)
    A = rng.integers(-128, 128, (5, 30), dtype=np.int8)
    Bd = rng.integers(-128, 128, (30, 45), dtype=np.int8)
    Bd[14:28, 0:14] = 0
    Cc = np.empty((5, 45), np.int32)
    nnz = R.ref_pack_and_bsr_matmul_int8(A, Bd, Cc, 5, 30, 45)
    rp = np.zeros(4, np.int64); ci = np.zeros(12, np.int64); dat = np.zeros(12 * 196, np.int8)
    nnz2 = R.ref_pack_to_bsr(Bd, 30, 45, rp, ci, dat)
    assert nnz == nnz2 == 11
    cc["convA_A"], cc["convA_B"], cc["convA_C"] = A, Bd, Cc
    cc["convA_row_ptr"], cc["convA_col_idx"], cc["convA_data"] = rp, ci[:nnz], dat[:nnz * 196]
    np.savez_compressed(os.path.join(OUT, "cpp_golden_cases.npz"), **cc)

    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"generator": "tools/make_golden.py", "numpy": np.__version__, **manifest,
                   "files": sorted(x for x in os.listdir(OUT) if x.endswith(".npz"))}, f, indent=1)
    for x in sorted(os.listdir(OUT)):
        print(f"{x:32s} {os.path.getsize(os.path.join(OUT, x)):9d} B")


if __name__ == "__main__":
    main()
