"""BASELINE config 1: MNIST-CNN INT8 with a BSR-pruned FC1 (14x14 blocks), batch 64, GPU path vs the CPU oracle chain.

Weights, scales and biases are the reference's ``data/int8`` files; the 32 golden images and FP32 logits are
``sw/golden/mnist_inputs.npy`` / ``mnist_logits_fp32.npy`` (all inside tests/golden/mnist_int8.npz, written by
tools/make_golden.py).  The batch is the 32 golden images + 32 seeded random images.
"""
import numpy as np
import pytest

from oracle import bsr_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN_LABELS = [7, 2, 1, 0, 4, 1, 4, 9, 6, 9, 0, 6, 9, 0, 1, 5, 9, 7, 8, 4, 9, 6, 6, 5, 4, 0, 7, 4, 0, 1, 3, 1]


def _batch64(w):
    rng = np.random.default_rng(64)
    return np.concatenate([w["inputs_u8"], rng.integers(0, 256, (32, 28, 28), dtype=np.uint8)], axis=0)


@pytest.mark.parametrize("mode,sparsity", [("raw", 0.0), ("raw", 0.9), ("normalized", 0.9), ("normalized", 0.5)])
def test_mnist_cnn_matches_oracle(golden, mode, sparsity):
    import torch
    from resnet_accel_b200.mnist import MnistCnnInt8
    w = dict(golden("mnist_int8.npz"))
    assert w["logits_fp32"].argmax(1).tolist() == GOLDEN_LABELS
    imgs = _batch64(w)
    net = MnistCnnInt8(w, batch=64, fc1_sparsity=sparsity, mode=mode)
    scales = net.calibrate(imgs)
    # the calibrator against the CPU restatement of the same float forward (float32 accumulation order differs: tolerance)
    ref_scales = O.mnist_activation_scales(O.mnist_float_forward(O.mnist_preprocess(imgs, mode), w))
    for k, v in ref_scales.items():
        assert abs(scales[k] - v) <= 2e-5 * v, (k, scales[k], v)
    net.build()
    # GPU pruner + packer == oracle pruner + packer
    fc1_ref = O.mnist_prune_fc1(w["fc1_weight_int8"], sparsity) if sparsity > 0 else O.build_bsr_14x14_int8_direct(w["fc1_weight_int8"])
    for key in ("indptr", "indices", "data"):
        assert np.array_equal(net.bsr["fc1"][key].cpu().numpy(), fc1_ref[key]), key
    assert fc1_ref["num_blocks"] == 6590 - int(6590 * sparsity)
    logits = net.run(imgs).cpu().numpy().copy()
    net.sat.zero_()                                           # the first call also ran the warm-up forward of the capture
    again = net.run(imgs).cpu().numpy()                       # graph replay
    assert np.array_equal(logits, again)
    ref = O.mnist_cnn_int8_forward(imgs, w, scales, fc1_ref, mode)        # same scales on both sides: bit-exact from here on
    assert np.array_equal(net.x_q.cpu().numpy(), ref["input_q"])
    for name, key in (("conv1", "conv1_out"), ("conv2", "conv2_out"), ("pool", "pooled"), ("fc1", "fc1_out"), ("logits_i32", "logits_i32")):
        assert np.array_equal(net.buf[name].cpu().numpy(), ref[key]), name
    for n in ("conv1", "conv2", "fc1", "fc2"):
        assert np.array_equal(net.bias[n].cpu().numpy(), ref["bias_i32"][n]), n
    assert np.array_equal(logits, ref["logits"])
    assert int(net.sat.item()) == ref["sat_count"]
    pred = logits.argmax(1)
    hits = int((pred[:32] == np.array(GOLDEN_LABELS)).sum())
    if mode == "raw" and sparsity == 0.0:
        # the only end-to-end anchor the reference ships: FP32 logits of these 32 images on raw pixels
        assert hits == 32
        assert np.abs(logits[:32] - w["logits_fp32"]).max() < 40.0      # logits of magnitude ~2000
    else:
        assert hits >= 28, hits                                # pruned without fine-tuning: 29-31 of 32 (oracle says the same)
    assert hits == int((ref["pred"][:32] == np.array(GOLDEN_LABELS)).sum())


def test_mnist_int8_dir_loader(tmp_path, golden):
    """load_int8_dir reads the reference's data/int8 layout (quantize.py:185-208)."""
    import json
    from resnet_accel_b200.mnist import load_int8_dir
    w = golden("mnist_int8.npz")
    for n in ("conv1", "conv2", "fc1", "fc2"):
        np.save(tmp_path / f"{n}_weight_int8.npy", w[f"{n}_weight_int8"])
        np.save(tmp_path / f"{n}_weight_scales.npy", w[f"{n}_weight_scales"])
        np.save(tmp_path / f"{n}_bias_int8.npy", w[f"{n}_bias_int8"])
        with open(tmp_path / f"{n}_bias_scale.json", "w") as f:
            json.dump({"scale": float(w[f"{n}_bias_scale"])}, f)
    got = load_int8_dir(str(tmp_path))
    for k in got:
        assert np.array_equal(np.asarray(got[k]), np.asarray(w[k])), k
