"""GPU halves of the round-2 boundary rows: the generic / ResNet INT8 file writers (device quantisers), relu / relu6 / relu32
kernels, the FP32 compatibility golden ``gemm_bsr_int8``, the pruning schedule and the whole-model quantiser - against
fixtures produced by the reference's own functions (tests/golden/r2_cases.npz, golden_fp32_cases.npz)."""
import numpy as np
import pytest

from oracle import bsr_oracle as O

pytestmark = pytest.mark.gpu


def test_generic_and_resnet_int8_writers(golden, tmp_path):
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    for tag in ("g32", "g64"):
        b = int(g[f"{tag}_block"])
        bsr = E.build_bsr_from_dense(g[f"{tag}_w"], b, b)                        # GPU packer (float data keeps its values)
        ref = O.build_bsr_from_dense(g[f"{tag}_w"], b, b)
        assert np.array_equal(bsr["indptr"], ref["indptr"]) and np.array_equal(bsr["indices"], ref["indices"])
        bsr["data"] = ref["data"]                                               # float64 blocks: the packer works in float32
        for suffix, sc in (("int8_bytes", g[f"{tag}_scales"]), ("int8_short_bytes", g[f"{tag}_scales"][:2]),
                           ("int8_scalar_bytes", np.asarray(0.013, dtype=g[f"{tag}_scales"].dtype))):
            p = tmp_path / f"{tag}_{suffix}.bsr"
            E.save_bsr_binary_int8_generic(bsr, sc, str(p))
            assert p.read_bytes() == g[f"{tag}_{suffix}"].tobytes(), (tag, suffix)
    # ResNet exporter: headered blob with per-channel scales inside the blocks
    bsr = E.build_bsr_from_dense(g["rn_w"], 14, 14)
    q, sc = E.quantize_symmetric_per_channel(g["rn_w"])
    assert np.array_equal(q, g["rn_q"]) and np.array_equal(sc, g["rn_scales"])
    p = tmp_path / "rn.bsr"
    E.save_bsr_binary_int8_resnet(bsr, sc, str(p))
    assert p.read_bytes() == g["rn_bytes"].tobytes()
    E.save_bsr_binary_int8_resnet(bsr, sc[:20], str(p))
    assert p.read_bytes() == g["rn_short_bytes"].tobytes()
    # reader of the same blob: structure and int8 blocks back
    d = E.deserialize_from_hardware(g["rn_bytes"].tobytes())
    assert np.array_equal(d["indptr"], bsr["indptr"]) and np.array_equal(d["data"], E.quantize_blocks_per_channel(bsr, sc))
    big = dict(bsr)
    big["indptr"] = np.array([0, 70000], np.int64)
    with pytest.raises(ValueError):
        E.save_bsr_binary_int8_resnet(big, sc, str(p))


def test_serialize_for_hardware_matches_reference_blob(golden):
    """bsr_packer.hpp:489-525: the GPU packer's structure + blocks, serialised, equal the bytes the reference wrote."""
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    bsr = E.build_bsr_14x14_int8_direct(g["hw_dense"])
    blob = E.serialize_for_hardware(bsr)
    assert blob == g["hw_blob"].tobytes()
    back = E.deserialize_from_hardware(blob)
    for k in ("indptr", "indices", "data"):
        assert np.array_equal(back[k], bsr[k]), k


def test_relu_kernels(golden):
    import torch
    from resnet_accel_b200 import ops
    g = golden("r2_cases.npz")
    x = np.arange(-128, 128, dtype=np.int8)
    for i in range(5):
        t = torch.from_numpy(np.tile(x, 37)).cuda()
        ops.relu6_int8(t, float(g[f"relu6_{i}_scale"]))
        assert np.array_equal(t.cpu().numpy(), np.tile(g[f"relu6_{i}_out"], 37)), i
    t = torch.from_numpy(x.copy()).cuda()
    assert np.array_equal(ops.relu_int8(t).cpu().numpy(), np.maximum(x, 0))
    t = torch.from_numpy(g["relu32_in"]).cuda()
    assert np.array_equal(ops.relu_int32(t).cpu().numpy(), g["relu32_out"])


def test_gemm_bsr_int8_fp32_compat(golden):
    """sw/golden/gemm_bsr_int8.py:16-104 replayed on the GPU, bit for bit against the reference function's own outputs."""
    import torch
    from resnet_accel_b200 import golden as G
    g = golden("golden_fp32_cases.npz")
    for i in range(3):
        b = int(g[f"f{i}_block"])
        bsr = O.build_bsr_from_dense(g[f"f{i}_B"], b, b)
        sA = g[f"f{i}_scaleA"][()]
        if sA.dtype == np.float64:      # generated from a Python float: weak scalar under NEP 50
            sA = float(sA)
        C = G.gemm_bsr_int8(g[f"f{i}_A"], bsr, sA, g[f"f{i}_scales"])
        assert isinstance(C, np.ndarray) and C.dtype == np.float32 and C.shape == g[f"f{i}_C"].shape
        assert np.array_equal(C, g[f"f{i}_C"]), i
        Ct = G.gemm_bsr_int8(torch.from_numpy(g[f"f{i}_A"]).cuda(), bsr, sA, g[f"f{i}_scales"])
        assert Ct.is_cuda and np.array_equal(Ct.cpu().numpy(), g[f"f{i}_C"])
    # beyond the fixtures: random shapes against the oracle's replay, float32 and float64 scales, strong / weak scale_A
    rng = np.random.default_rng(3)
    for b, K, N, M, sdt, strong in ((8, 64, 40, 7, np.float32, False), (14, 42, 70, 5, np.float64, False),
                                    (4, 32, 30, 9, np.float32, True), (8, 48, 64, 3, np.float64, True)):
        B = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
        keep = rng.random((-(-K // b), -(-N // b))) < 0.6
        B = B * np.repeat(np.repeat(keep, b, 0), b, 1)[:K, :N]
        bsr = O.build_bsr_from_dense(B, b, b)
        A = rng.integers(-128, 128, (M, K), dtype=np.int8)
        scales = rng.uniform(0.001, 0.01, K).astype(sdt)
        sA = np.float64(0.0123) if strong else 0.0123
        want = O.gemm_bsr_int8_fp32(A, bsr, sA, scales)
        got = G.gemm_bsr_int8(A, bsr, sA, scales)
        assert np.array_equal(got, want), (b, K, N, sdt, strong)
    with pytest.raises(ValueError):
        G.gemm_bsr_int8(np.zeros((2, 30), np.int8), O.build_bsr_from_dense(np.ones((30, 16), np.float32), 8, 8), 1.0,
                        np.ones(30, np.float32))


def test_progressive_sparsity_schedule(golden):
    """blocksparse_train.py:282-321 without the fine-tuning: masks and masked weights after each of the four phases."""
    import torch
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    names = ("conv1", "fc1", "fc2")
    weights = {n: torch.from_numpy(g[f"sched_{n}_w"].copy()).cuda() for n in names}
    masks = {n: torch.ones_like(weights[n], dtype=torch.bool) for n in names}
    seen = []
    hist = E.progressive_sparsity(weights, masks, 0.9,
                                  fine_tune=lambda w, m, s: seen.append((s, {n: w[n].detach().cpu().numpy().copy() for n in names},
                                                                         {n: m[n].detach().cpu().numpy().copy() for n in names})))
    assert [h["sparsity"] for h in hist] == [0.5, 0.7, 0.85, 0.9]
    for phase, (s, w, m) in enumerate(seen):
        assert hist[phase]["blocks_pruned"] == int(g[f"sched_{phase}_pruned"]) or phase > 0
        for n in names:
            assert np.array_equal(m[n].reshape(m[n].shape[0], -1), g[f"sched_{phase}_{n}_mask"]), (phase, n)
            assert np.array_equal(w[n].reshape(w[n].shape[0], -1), g[f"sched_{phase}_{n}_w"]), (phase, n)
    total_pruned = sum(int((~seen[-1][2][n]).sum()) for n in names)
    assert total_pruned > 0


def test_quantize_model_per_channel(golden):
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    params = {}
    for n in ("conv1", "fc1", "fc2"):
        params[f"{n}.weight"] = g[f"qm_{n}_weight_fp32"]
        params[f"{n}.bias"] = g[f"qm_{n}_bias_fp32"]
    qm = E.quantize_model_per_channel(params)
    for n in ("conv1", "fc1", "fc2"):
        rec = qm[f"{n}.weight"]
        assert np.array_equal(rec["data"], g[f"qm_{n}_weight_data"]) and np.array_equal(rec["scales"], g[f"qm_{n}_weight_scales"])
        assert rec["axis"] == 0 and tuple(rec["shape"]) == g[f"qm_{n}_weight_fp32"].shape
        want = g[f"qm_{n}_weight_err"]
        got = np.array([rec["error"][e] for e in ("max_error", "mean_error", "mse", "snr_db")])
        assert np.allclose(got, want, rtol=1e-5, atol=1e-9)
        rb = qm[f"{n}.bias"]
        assert np.array_equal(rb["data"], g[f"qm_{n}_bias_data"]) and rb["scale"] == float(g[f"qm_{n}_bias_scale"])
