"""GPU parity of the whole ResNet-18 BSR network (CUDA graph replay) against the reference C++ golden chain."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.mark.parametrize("sparsity", [70.0, 90.0])
def test_resnet18_full_network_matches_cpu_golden(sparsity):
    """BASELINE configs 3 (70 %) and 4 (90 %): the whole 224x224 network, bit for bit against the reference C++ golden chain."""
    import torch
    import bench
    from oracle import bsr_oracle as O
    from oracle import c_oracle
    from resnet_accel_b200 import layers as L
    B = 2
    net = L.BsrNetwork(L.resnet18_specs(), sparsity, B)
    x = np.random.default_rng(0).integers(-128, 128, (B, 3, 224, 224), dtype=np.int8)
    xd = torch.from_numpy(x).cuda()
    eager = net.forward(xd).cpu().numpy().copy()
    net.capture(xd)
    replay = net.replay().cpu().numpy()
    assert np.array_equal(eager, replay)
    specs, hl, s_out = bench.host_network(sparsity)
    # GPU quantiser/packer produced the same INT8 weights as the host recipe
    for name, lay in net.layers.items():
        dense = O.bsr_to_dense(lay.bsr["indptr"].cpu().numpy(), lay.bsr["indices"].cpu().numpy(),
                               lay.bsr["data"].cpu().numpy(), lay.bsr["num_block_cols"])
        q2 = hl[name]["q"].reshape(hl[name]["q"].shape[0], -1)
        assert np.array_equal(dense[:q2.shape[0], :q2.shape[1]], q2), name
    if c_oracle.have_ref():
        for b in range(B):
            ref = bench.cpu_forward_image(x[b], specs, hl, s_out)
            assert np.array_equal(replay[b], ref.reshape(-1)), b
    else:  # port: layer by layer with the C restatement
        pytest.skip("oracle/_ref not available")
    if sparsity <= 70.0:
        assert replay.any()
    else:
        # with 10 % of the weights and the fixed synthetic scales the signal dies out before the logits (GPU and reference
        # agree on that, bit for bit); the early tensors still carry data, and test_layers_at_90pct_with_random_inputs below
        # exercises every layer of the 90 % network on full-range inputs
        assert net.buffers["layer1.0.conv1"].any().item()


def test_layers_at_90pct_with_random_inputs():
    """BASELINE config 4 (90 % block sparsity): every convolution / FC of the 224x224 ResNet-18, each fed its own full-range
    random int8 input (and residual), against the plain-C port.  Batch 2."""
    import torch
    from oracle import bsr_oracle as O
    from oracle import c_oracle
    from resnet_accel_b200 import layers as L, ops
    B = 2
    specs = L.resnet18_specs()
    net = L.BsrNetwork(specs, 90.0, B, bias_range=500)
    rng = np.random.default_rng(90)
    for sp in specs:
        if sp.kind not in ("conv", "fc"):
            continue
        lay = net.layers[sp.name]
        bsr = {k: lay.bsr[k].cpu().numpy() for k in ("indptr", "indices", "data")}
        total = lay.plan.n_block_rows * lay.plan.n_block_cols
        assert lay.plan.num_blocks == total - int(total * 0.9), sp.name
        bias, sf = lay.bias.cpu().numpy(), lay.sf.cpu().numpy()
        if sp.kind == "fc":
            A = rng.integers(-128, 128, (B, sp.c_in), dtype=np.int8)
            got = lay.plan.gemm(torch.from_numpy(A).cuda(), "i32", n_channels=sp.c_out, bias=lay.bias).cpu().numpy()
            want, _ = O.linear_bsr_layer(A, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, bias=bias)
            assert np.array_equal(got, want), sp.name
            continue
        x = rng.integers(-128, 128, (B, sp.c_in, sp.h, sp.w), dtype=np.int8)
        xd = ops.alloc_padded(x.shape); xd.copy_(torch.from_numpy(x).cuda())
        out = ops.alloc_padded((B, sp.c_out, sp.h_out, sp.w_out))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        if sp.residual:
            r = rng.integers(-128, 128, (B, sp.c_out, sp.h_out, sp.w_out), dtype=np.int8)
            rd = ops.alloc_padded(r.shape); rd.copy_(torch.from_numpy(r).cuda())
            got = lay.plan.conv(xd, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, bias=lay.bias, relu=False,
                                residual=rd, res_scales=(0.05, 0.05, 0.05), relu_out=True, out=out, sat_count=cnt).cpu().numpy()
            want, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, sp.k, sp.stride, sp.pad,
                                                bias=bias, relu=False, sf=sf, residual=r, res_scales=(0.05, 0.05, 0.05))
            want = np.maximum(want, 0)
        else:
            got = lay.plan.conv(xd, sp.k, sp.stride, sp.pad, sp.c_out, "i8", chan_scale=lay.sf, bias=lay.bias, relu=sp.relu,
                                out=out, sat_count=cnt).cpu().numpy()
            want, sat = c_oracle.conv_bsr_layer(x, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, sp.k, sp.stride, sp.pad,
                                                bias=bias, relu=sp.relu, sf=sf)
        assert np.array_equal(got, want), sp.name
        assert int(cnt.item()) == sat, sp.name
        assert got.any(), sp.name


def _check_layers_against_port(net, specs, x, s_in_expected, res_scales_expected):
    """Every conv layer of ``net`` against the plain-C port; the requant factors and residual scales are recomputed here
    from the expected per-tensor scales (not read back from the network)."""
    import torch
    from oracle import bsr_oracle as O
    from oracle import c_oracle
    net.forward(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    t = {"input": x}
    prev = "input"
    for sp in specs:
        src = t[sp.src] if sp.src else t[prev]
        if sp.kind in ("conv", "fc"):
            lay = net.layers[sp.name]
            sf = O.channel_scale_factors(s_in_expected[sp.name], lay.w_scales, lay.s_out)
            assert np.array_equal(lay.sf.cpu().numpy(), sf), sp.name
        if sp.name in net.fused_pool:       # stem convolution fused with its max-pool: its own output is never materialised
            bsr = {k: lay.bsr[k].cpu().numpy() for k in ("indptr", "indices", "data")}
            t[sp.name], _ = c_oracle.conv_bsr_layer(src, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, sp.k, sp.stride,
                                                    sp.pad, bias=None if lay.bias is None else lay.bias.cpu().numpy(),
                                                    relu=sp.relu, sf=sf)
            prev = sp.name
            continue
        got = net.buffers[sp.name].cpu().numpy()
        if sp.kind == "maxpool" and prev in net.fused_pool:
            want = np.stack([O.maxpool2d_int8(src[b], sp.k, sp.stride, sp.pad) for b in range(src.shape[0])])
            assert np.array_equal(got, want), sp.name
        if sp.kind == "conv":
            bsr = {k: lay.bsr[k].cpu().numpy() for k in ("indptr", "indices", "data")}
            bias = None if lay.bias is None else lay.bias.cpu().numpy()
            if sp.residual:
                want, _ = c_oracle.conv_bsr_layer(src, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, sp.k, sp.stride,
                                                  sp.pad, bias=bias, relu=False, sf=sf, residual=t[sp.residual],
                                                  res_scales=res_scales_expected[sp.name])
                want = np.maximum(want, 0)
            else:
                want, _ = c_oracle.conv_bsr_layer(src, bsr["indptr"], bsr["indices"], bsr["data"], sp.c_out, sp.k, sp.stride,
                                                  sp.pad, bias=bias, relu=sp.relu, sf=sf)
            assert np.array_equal(got, want), sp.name
        t[sp.name] = got            # continue from the device result so that one mismatch does not cascade
        prev = sp.name


def test_network_layers_match_oracle_port():
    """Each conv layer of a down-scaled ResNet-18 (64x64 images) against the plain-C port, incl. residual + ReLU.
    Synthetic recipe: every layer reads at S_ACT_IN and writes at S_ACT_OUT."""
    from resnet_accel_b200 import layers as L
    B = 3
    specs = L.resnet18_specs(image=64, num_classes=100)
    net = L.BsrNetwork(specs, 70.0, B, bias_range=300)
    x = np.random.default_rng(1).integers(-128, 128, (B, 3, 64, 64), dtype=np.int8)
    s_in = {sp.name: L.S_ACT_IN for sp in specs}
    res = {sp.name: (L.S_ACT_OUT, L.S_ACT_OUT, L.S_ACT_OUT) for sp in specs}
    _check_layers_against_port(net, specs, x, s_in, res)


def test_chained_scales_differ_from_module_constants():
    """scales chain through the network.  Input quantised at 0.031, every convolution writes at 0.043: a layer
    reads at its producer's output scale, the residual add uses (own output, identity tensor, own output)."""
    from resnet_accel_b200 import layers as L
    B, s_input, s_out = 2, 0.031, 0.043
    specs = L.resnet18_specs(image=64, num_classes=100)
    net = L.BsrNetwork(specs, 70.0, B, bias_range=200, s_input=s_input, s_out=s_out, chain_scales=True)
    x = np.random.default_rng(2).integers(-128, 128, (B, 3, 64, 64), dtype=np.int8)
    s_in, res = {}, {}
    for sp in specs:
        if sp.kind in ("conv", "fc"):
            s_in[sp.name] = s_input if sp.name == "conv1" else s_out        # pools keep the scale, so everything else reads s_out
            res[sp.name] = (s_out, s_out, s_out)
    assert net.scale_of["maxpool"] == s_out and net.scale_of["input"] == s_input
    _check_layers_against_port(net, specs, x, s_in, res)
    # an engine loaded from files chains the same way: its layer1 convolutions read at s_out, not at s_in
    eng_scales = L.input_scales(specs, s_input, s_out)
    assert eng_scales["conv1"] == s_input and eng_scales["layer1.0.conv1"] == s_out and eng_scales["fc"] == s_out


def test_resnet_inference_engine_roundtrip(tmp_path):
    """ResNetInference.load_model on the directory layout of resnet_inference.hpp reproduces the network it was saved from."""
    import torch
    from oracle import bsr_oracle as O
    from resnet_accel_b200 import layers as L
    B = 2
    specs = L.resnet18_specs(image=64, num_classes=50)
    ref = L.BsrNetwork(specs, 70.0, B, chain_scales=True)          # the engine chains scales: so must the network it is saved from
    for name, lay in ref.layers.items():
        dense = O.bsr_to_dense(lay.bsr["indptr"].cpu().numpy(), lay.bsr["indices"].cpu().numpy(), lay.bsr["data"].cpu().numpy(),
                               lay.bsr["num_block_cols"])
        sp = lay.spec
        np.save(tmp_path / f"{name}_weight_int8.npy", dense[:sp.c_out, :sp.K].astype(np.int8).reshape(sp.c_out, sp.c_in, sp.k, sp.k))
        np.save(tmp_path / f"{name}_weight_scales.npy", lay.w_scales)
    eng = L.ResNetInference(batch=B, image=64, num_classes=50)
    with pytest.raises(RuntimeError):
        eng.run_inference(torch.zeros((B, 3, 64, 64), dtype=torch.int8))
    eng.load_model(str(tmp_path))
    x = torch.randint(-128, 128, (B, 3, 64, 64), dtype=torch.int8, device="cuda")
    want = ref.forward(x).cpu().numpy().copy()
    got = eng.run_inference(x).cpu().numpy()
    assert np.array_equal(got, want)
    assert 0.6 < eng.get_model_sparsity() < 0.8
    idx, prob = eng.get_top_k(torch.from_numpy(got).cuda(), k=3)
    assert idx.shape == (B, 3) and np.all(prob[:, 0] >= prob[:, 1])
    assert eng.benchmark(3)["images_per_s"] > 0


def test_pipelined_serving_loop_matches_single_batches():
    """run_inference_pipelined (two alternating graphs, copies on side streams) returns, for every host batch, the logits
    run_inference gives for that batch alone - also when the ring of pinned buffers is shorter than the number of batches in
    flight allows (each buffer reused every second step) and when the loop is called twice."""
    import torch
    from resnet_accel_b200 import layers as L
    B, n = 4, 5
    eng = L.ResNetInference(batch=B, image=64, num_classes=50)
    eng.load_synthetic(70.0, chain_scales=True)
    g = torch.Generator().manual_seed(5)
    xs = [torch.randint(-128, 128, (B, 3, 64, 64), dtype=torch.int8, generator=g).pin_memory() for _ in range(n)]
    want = [eng.run_inference(x).cpu().numpy().copy() for x in xs]
    assert any(not np.array_equal(want[0], w) for w in want[1:])          # the batches really differ
    for _ in range(2):
        ys = [torch.zeros((B, 50), dtype=torch.int32).pin_memory() for _ in range(n)]
        eng.run_inference_pipelined(xs, ys)
        torch.cuda.current_stream().synchronize()
        for i in range(n):
            assert np.array_equal(ys[i].numpy(), want[i]), i
    # a ring of two pinned buffers, refilled by the host between calls (as bench.py does)
    ring_x = [xs[0].clone().pin_memory(), xs[1].clone().pin_memory()]
    ring_y = [torch.zeros((B, 50), dtype=torch.int32).pin_memory() for _ in range(2)]
    eng.run_inference_pipelined([ring_x[i & 1] for i in range(4)], [ring_y[i & 1] for i in range(4)])
    torch.cuda.current_stream().synchronize()
    assert np.array_equal(ring_y[0].numpy(), want[0]) and np.array_equal(ring_y[1].numpy(), want[1])
