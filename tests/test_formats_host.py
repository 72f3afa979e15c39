"""Host-side halves of the file formats and mask recipes (no GPU needed): product code against fixtures produced by the
reference's own functions (tools/make_golden_r2.py -> tests/golden/r2_cases.npz)."""
import json

import numpy as np
import pytest

from oracle import bsr_oracle as O


def test_deserialize_from_hardware_matches_reference_reader(golden):
    """bsr_packer.hpp:530-575 (reader) on the blob the reference's serialize_for_hardware wrote (:489-525)."""
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    blob = g["hw_blob"].tobytes()
    d = E.deserialize_from_hardware(blob)
    assert [d["num_blocks"], d["num_block_rows"], d["num_block_cols"]] == g["hw_hdr"].tolist()
    assert np.array_equal(d["indptr"], g["hw_row_ptr"]) and np.array_equal(d["indices"], g["hw_col_idx"])
    assert np.array_equal(d["data"].reshape(-1), g["hw_data"])
    # the packer's structure: blocks of the dense matrix with any non-zero element
    ref = O.build_bsr_14x14_int8_direct(g["hw_dense"])
    assert np.array_equal(ref["indptr"], d["indptr"]) and np.array_equal(ref["data"], d["data"])
    assert O.serialize_for_hardware(ref["indptr"], ref["indices"], ref["data"], ref["num_block_cols"]) == blob
    with pytest.raises(ValueError, match="Buffer too small for BSR header"):
        E.deserialize_from_hardware(blob[:11])
    with pytest.raises(ValueError, match="Buffer size mismatch: expected"):
        E.deserialize_from_hardware(blob[:-1])


def test_create_block_sparse_mask_replays_torch_rng(golden):
    """train_resnet18.py:69-132: torch.manual_seed + randperm on the CPU generator."""
    import torch
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    for i in range(3):
        cfg = g[f"bmask_{i}_cfg"].tolist()
        shape = tuple(v for v in cfg[:4] if v > 0)
        bs, sp, seed = (cfg[4], cfg[5]), cfg[6] / 100.0, cfg[7]
        m = E.create_block_sparse_mask(torch.zeros(shape), bs, sp, seed)
        assert tuple(m.shape) == shape and m.dtype == torch.float32
        assert np.array_equal(m.numpy(), g[f"bmask_{i}"]), i


def test_layer_tables_and_metadata_writers(golden, tmp_path):
    from resnet_accel_b200 import exporters as E
    g = golden("r2_cases.npz")
    want = json.loads(g["rn_layer_config_json"].tobytes().decode())
    got = json.loads(json.dumps(E.get_resnet18_layer_config(), sort_keys=True))
    assert got == want
    cfg = json.loads(g["sched_cfg_json"].tobytes().decode())
    for n, w in (("conv1", np.zeros((12, 3, 3, 3))), ("fc1", np.zeros((40, 64))), ("fc2", np.zeros((16, 40)))):
        (bh, bw), keep = E.layer_block_cfg(n, w)
        assert [[bh, bw], keep] == [list(cfg[n][0]), cfg[n][1]]
    # generic metadata (export_bsr.py:205-241) and ResNet metadata (export_resnet18_bsr.py:250-266): same JSON documents
    for tag in ("g32", "g64"):
        b = int(g[f"{tag}_block"])
        bsr = O.build_bsr_from_dense(g[f"{tag}_w"], b, b)
        p = tmp_path / f"{tag}.json"
        E.save_bsr_metadata_generic(bsr, str(p), layer_name=tag)
        assert json.loads(p.read_text()) == json.loads(g[f"{tag}_meta_json"].tobytes().decode())
        E.save_bsr_binary(bsr, str(tmp_path / "f32.bsr"))
        assert (tmp_path / "f32.bsr").read_bytes() == g[f"{tag}_f32_bytes"].tobytes()
    bsr = O.build_bsr_from_dense(g["rn_w"], 14, 14)
    p = tmp_path / "rn.json"
    E.save_bsr_metadata_resnet(bsr, str(p), layer_name="rn")
    assert json.loads(p.read_text()) == json.loads(g["rn_meta_json"].tobytes().decode())


def test_oracle_restatements_of_round2_rows(golden):
    """The CPU restatements the GPU tests compare against, pinned on the reference's outputs."""
    g = golden("r2_cases.npz")
    for tag in ("g32", "g64"):
        b = int(g[f"{tag}_block"])
        bsr = O.build_bsr_from_dense(g[f"{tag}_w"], b, b)
        assert O.quantize_blocks_by_block_row(bsr, g[f"{tag}_scales"]).tobytes() == g[f"{tag}_int8_bytes"].tobytes()
        assert O.quantize_blocks_by_block_row(bsr, g[f"{tag}_scales"][:2]).tobytes() == g[f"{tag}_int8_short_bytes"].tobytes()
    bsr = O.build_bsr_from_dense(g["rn_w"], 14, 14)
    assert O.serialize_resnet_layer(bsr, g["rn_scales"]) == g["rn_bytes"].tobytes()
    assert O.serialize_resnet_layer(bsr, g["rn_scales"][:20]) == g["rn_short_bytes"].tobytes()
    x = np.arange(-128, 128, dtype=np.int8)
    for i in range(5):
        assert np.array_equal(O.relu6_int8(x, float(g[f"relu6_{i}_scale"])), g[f"relu6_{i}_out"]), i
    assert np.array_equal(O.relu_int32(g["relu32_in"]), g["relu32_out"])
    # pruning schedule: four phases of the global selection on the masked weights (no fine-tuning in between)
    names = ("conv1", "fc1", "fc2")
    w = {n: g[f"sched_{n}_w"].reshape(g[f"sched_{n}_w"].shape[0], -1).copy() for n in names}
    cfg = {"conv1": ((4, 4), 0.30), "fc1": ((8, 8), 0.05), "fc2": ((8, 8), 0.05)}
    masks = {n: np.ones(w[n].shape, bool) for n in names}
    for phase, sp in enumerate((0.5, 0.7, 0.85, 0.9)):
        norms = [O.compute_block_norms(w[n], *cfg[n][0]) for n in names]
        keeps = O.prune_blocks_global(norms, sp, [cfg[n][1] for n in names])
        pruned = 0
        for n, k in zip(names, keeps):
            (bh, bw) = cfg[n][0]
            full = np.repeat(np.repeat(k, bh, 0), bw, 1)[:w[n].shape[0], :w[n].shape[1]]
            masks[n] &= full
            w[n] = w[n] * masks[n]
            pruned += int((~k).sum())
            assert np.array_equal(masks[n], g[f"sched_{phase}_{n}_mask"]), (phase, n)
            assert np.array_equal(w[n], g[f"sched_{phase}_{n}_w"]), (phase, n)
        assert pruned == int(g[f"sched_{phase}_pruned"])


def test_padded200_block_file_roundtrip(golden, tmp_path):
    """The 200-byte block variant of export_bsr_14x14.py:17-21 next to the shipped 196-byte layout."""
    from resnet_accel_b200 import exporters as E
    mn = golden("mnist_int8.npz")
    bsr = O.build_bsr_14x14_int8_direct(mn["fc2_weight_int8"])
    p196, p200 = tmp_path / "w196.bsr", tmp_path / "w200.bsr"
    E.save_bsr_binary_int8(bsr, str(p196))
    E.save_bsr_binary_int8_padded200(bsr, str(p200))
    assert p196.stat().st_size == bsr["num_blocks"] * 196 and p200.stat().st_size == bsr["num_blocks"] * 200
    assert p200.stat().st_size % 8 == 0
    assert np.array_equal(E.load_bsr_binary_int8(str(p196)), bsr["data"])
    assert np.array_equal(E.load_bsr_binary_int8(str(p200), padded=True), bsr["data"])
    with pytest.raises(ValueError):
        E.load_bsr_binary_int8(str(p196), padded=True) if (bsr["num_blocks"] * 196) % 200 else (_ for _ in ()).throw(ValueError())
