"""Pins the CPU oracle (oracle/*.py) to the reference's own behaviour.

Fixtures in tests/golden/ were produced by the REAL reference (oracle/gen_golden.py).  In the
build container the real reference module is additionally executed side by side.
"""
import os
import warnings

import pytest
import torch

from oracle import cunet_oracle, quantize_oracle, evaluation_oracle, ref_loader, synthetic

TINY = ["tiny_L3_K2.pt", "tiny_L2_K1.pt", "tiny_L3_K0.pt"]


def _close(a, b, tol=2e-5):
    a, b = a.double(), b.double()
    scale = max(b.abs().max().item(), 1e-12)
    return (a - b).abs().max().item() / scale <= tol


@pytest.mark.parametrize("name", TINY)
def test_oracle_matches_reference_fixture(golden_dir, name):
    fx = torch.load(os.path.join(golden_dir, name), weights_only=False)
    cfg = fx["config"]
    net = cunet_oracle.OracleCUNet(fx["state_before"], cfg["class_num"], cfg["layer_num"],
                                   cfg["order"], cfg["loss_num"], neck_size=cfg["neck_size"],
                                   growth_rate=cfg["growth_rate"], init_chan_num=cfg["init_chan_num"])
    spec_names = [n for n, _, _ in cunet_oracle.state_spec(
        cfg["class_num"], cfg["layer_num"], cfg["order"], cfg["neck_size"], cfg["growth_rate"],
        cfg["init_chan_num"])]
    assert spec_names == list(fx["state_before"].keys())        # names + registration order
    for (n, shape, _), v in zip(cunet_oracle.state_spec(
            cfg["class_num"], cfg["layer_num"], cfg["order"], cfg["neck_size"],
            cfg["growth_rate"], cfg["init_chan_num"]), fx["state_before"].values()):
        assert tuple(v.shape) == tuple(shape), n
    net.train()
    outs = net(fx["img"])
    loss = cunet_oracle.multi_loss_mse(outs, fx["heatmap"])
    loss.backward()
    assert len(outs) == len(fx["outputs_train"])
    for o, r in zip(outs, fx["outputs_train"]):
        assert _close(o.detach(), r)
    assert _close(loss.detach(), fx["loss_train"])
    for n in net.param_names:
        g = net.state[n].grad
        g = torch.zeros_like(net.state[n]) if g is None else g
        assert _close(g, fx["grads"][n], 2e-4), n
    # BN running statistics incl. the double update of checkpointed BNs
    for k, v in fx["state_after_train"].items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert _close(net.state[k], v, 1e-5), k
        if k.endswith("num_batches_tracked"):
            assert int(net.state[k]) == int(v), k
    net.eval()
    with torch.no_grad():
        outs = net(fx["img"])
    for o, r in zip(outs, fx["outputs_eval"]):
        assert _close(o, r)


def test_oracle_real_config_digest(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "real_L2_K1_C68_n1.pt"), weights_only=False)
    cfg = fx["config"]
    state = cunet_oracle.init_state(cfg["class_num"], cfg["layer_num"], cfg["order"], seed=cfg["seed"])
    assert list(state.keys()) == fx["state_keys"]
    assert [tuple(v.shape) for v in state.values()] == fx["state_shapes"]
    assert len(state) == 396
    net = cunet_oracle.OracleCUNet(state, cfg["class_num"], cfg["layer_num"], cfg["order"],
                                   cfg["loss_num"])
    img, hm = synthetic.make_inputs(cfg["n"], cfg["class_num"], seed=cfg["seed"])
    outs = net(img)
    loss = cunet_oracle.multi_loss_mse(outs, hm)
    loss.backward()
    assert _close(loss.detach(), fx["loss_train"], 1e-5)
    for o, r in zip(outs, fx["out_samples"]):
        assert _close(o.detach()[:, ::7, ::5, ::3], r, 1e-4)
    for k, r in fx["grad_samples"].items():
        assert _close(net.state[k].grad.flatten()[::97], r, 1e-3), k
    for k, r in fx["running_var_samples"].items():
        assert _close(net.state[k].flatten()[::13], r, 1e-5), k
    for k, r in fx["num_batches_tracked"].items():
        assert int(net.state[k]) == r, k


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only in build container")
def test_oracle_against_live_reference():
    torch.manual_seed(5)
    ref = ref_loader.create_reference_net(4, 3, 1, 3, neck_size=2, growth_rate=8, init_chan_num=16)
    net = cunet_oracle.OracleCUNet(ref.state_dict(), 4, 3, 1, 3, neck_size=2, growth_rate=8,
                                   init_chan_num=16)
    img = torch.rand(2, 3, 64, 64)
    hm = torch.rand(2, 4, 16, 16)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ro = ref(img)
        rl = cunet_oracle.multi_loss_mse(ro, hm)
        rl.backward()
    oo = net(img)
    ol = cunet_oracle.multi_loss_mse(oo, hm)
    ol.backward()
    assert _close(ol.detach(), rl.detach())
    for (k, p) in ref.named_parameters():
        assert _close(net.state[k].grad, p.grad, 2e-4), k


@pytest.mark.parametrize("bits_w", [1, 2, 8])
def test_quanop_oracle_matches_reference_fixture(golden_dir, bits_w):
    fx = torch.load(os.path.join(golden_dir, "quanop_bits%d.pt" % bits_w), weights_only=False)
    idx = quantize_oracle.target_indices(len(fx["w0"]))
    assert len(idx) == fx["num_targets"]
    tw = [fx["w0"][i] for i in idx]
    wq, saved = quantize_oracle.quanop_quantization(tw, bits_w, fx["bits_g"])
    for j, i in enumerate(idx):
        assert torch.equal(wq[j], fx["wq"][i]), (bits_w, i)
        assert torch.equal(saved[j], fx["wr"][i]), (bits_w, i)
    for i in range(len(fx["w0"])):                 # first / last conv untouched
        if i not in idx:
            assert torch.equal(fx["wq"][i], fx["w0"][i])
    gq = quantize_oracle.quanop_update_grad(saved, [fx["g0"][i] for i in idx], bits_w, fx["bits_g"])
    for j, i in enumerate(idx):
        assert torch.equal(gq[j], fx["gq"][i]), (bits_w, i)
    # value sets
    if bits_w in (1, 2):
        for w in wq:
            assert set(w.unique().tolist()) <= {-1.0, 0.0, 1.0}
    for g in gq:
        assert torch.equal(g * 128, torch.round(g * 128)) and g.abs().max() <= 0.9921875


def test_binop_properties():
    gen = torch.Generator().manual_seed(3)
    w = [(torch.rand(32, 128, 3, 3, generator=gen) * 2 - 1) * 1.3]
    wb, saved = quantize_oracle.binop_binarization(w)
    s = saved[0]
    assert s.abs().max() <= 1.0
    alpha = s.abs().mean(dim=(1, 2, 3), keepdim=True)
    assert torch.allclose(wb[0], s.sign() * alpha)
    g = [torch.randn(32, 128, 3, 3, generator=gen)]
    gu = quantize_oracle.binop_update_grad(saved, g)[0]
    n = 128 * 9
    expect = (alpha * (s.abs() <= 1).float() * g[0]
              + s.sign() * (s.sign() * g[0]).sum(dim=(1, 2, 3), keepdim=True) / n) * (1 - 1 / 128.) * n
    assert torch.allclose(gu, expect, rtol=1e-5, atol=1e-6)


def test_get_preds_semantics():
    s = torch.zeros(1, 3, 64, 64)
    s[0, 0, 10, 20] = 2.0
    s[0, 0, 30, 40] = 2.0        # tie -> first maximum
    s[0, 1] = -1.0               # max <= 0 -> masked
    s[0, 2, 63, 63] = 0.5
    p = evaluation_oracle.get_preds(s)
    assert p[0, 0].tolist() == [21.0, 11.0]
    assert p[0, 1].tolist() == [0.0, 0.0]
    assert p[0, 2].tolist() == [64.0, 64.0]
    img, hm = synthetic.make_inputs(2, 5, seed=1)
    p = evaluation_oracle.get_preds(hm)
    assert (p >= 5).all() and (p <= 60).all()


def test_product_synthetic_matches_oracle_copy():
    """cu-net.py / bench.py generate their batches with cunet_b200.utils.synthetic; the oracle keeps its own copy."""
    from cunet_b200.utils import synthetic as prod
    for seed in (0, 7):
        a, b = prod.make_inputs(2, 5, seed=seed), synthetic.make_inputs(2, 5, seed=seed)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
