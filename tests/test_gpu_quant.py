"""Quantizer kernels (QuanOp / BinOp / QuanInput) against the CPU oracle and the reference's own fixtures."""
import os

import pytest
import torch
import torch.nn as nn

from oracle import quantize_oracle, evaluation_oracle, synthetic

pytestmark = pytest.mark.gpu


def _model(ws):
    convs = []
    for w in ws:
        co, ci, k, _ = w.shape
        m = nn.Conv2d(ci, co, k, bias=False)
        m.weight.data = w.clone()
        convs.append(m)
    return nn.Sequential(*convs).cuda()


@pytest.mark.parametrize("bits_w", [1, 2, 8])
def test_quanop_matches_reference_fixture(golden_dir, bits_w):
    """Bit-exact against the outputs of the REAL utils/quantize.py (tests/golden/quanop_bits*.pt)."""
    from cunet_b200.utils.quantize import QuanOp
    fx = torch.load(os.path.join(golden_dir, "quanop_bits%d.pt" % bits_w), weights_only=False)
    model = _model(fx["w0"])
    op = QuanOp(model, bits_w=bits_w, bits_g=fx["bits_g"])
    assert op.num_of_params == fx["num_targets"]
    op.quantization()
    convs = [m for m in model.modules() if isinstance(m, nn.Conv2d)]
    for m, ref in zip(convs, fx["wq"]):
        assert torch.equal(m.weight.data.cpu(), ref)
    for m, g in zip(convs, fx["g0"]):
        m.weight.grad = g.clone().cuda()
    op.restore()
    for m, ref in zip(convs, fx["wr"]):
        assert torch.equal(m.weight.data.cpu(), ref)
    op.updateQuanGradWeight()
    for i, (m, ref) in enumerate(zip(convs, fx["gq"])):
        assert torch.equal(m.weight.grad.cpu(), ref), i


def test_binop_matches_reference_fixture(golden_dir):
    """BinOp kernels against the outputs of the REAL BinOp class (tests/golden/binop.pt)."""
    from cunet_b200.models.cu_net_prev_version import BinOp
    fx = torch.load(os.path.join(golden_dir, "binop.pt"), weights_only=False)
    model = _model(fx["w0"])
    op = BinOp(model)
    assert op.num_of_params == fx["num_targets"]
    op.binarization()
    convs = [m for m in model.modules() if isinstance(m, nn.Conv2d)]
    for i, (m, ref) in enumerate(zip(convs, fx["wb"])):
        got = m.weight.data.cpu()
        assert (got - ref).abs().max() <= 1e-6 * ref.abs().max(), i
        assert torch.equal(got.sign(), ref.sign()), i
    for m, g in zip(convs, fx["g0"]):
        m.weight.grad = g.clone().cuda()
    op.restore()
    for i, (m, ref) in enumerate(zip(convs, fx["wr"])):
        assert (m.weight.data.cpu() - ref).abs().max() <= 1e-7, i
    op.updateBinaryGradWeight()
    for i, (m, ref) in enumerate(zip(convs, fx["gb"])):
        assert (m.weight.grad.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max(), i


def test_quaninput_matches_reference_fixture(golden_dir):
    """QuanInput kernels against the REAL QuanInput.forward / .backward bodies (tests/golden/quaninput.pt): bit exact."""
    from cunet_b200.utils.quantize import QuanInput2d
    fx = torch.load(os.path.join(golden_dir, "quaninput.pt"), weights_only=False)
    for bits_i, d in fx.items():
        x = d["x"].cuda().requires_grad_(True)
        y = QuanInput2d(bits_i)(x)
        assert torch.equal(y.detach().cpu(), d["y"]), bits_i
        y.backward(d["gy"].cuda())
        assert torch.equal(x.grad.cpu(), d["gx"]), bits_i


def test_binop_target_set_on_the_dropin_model():
    """BinOp(net) on the CU-Net drop-in binarises the prev-version set (72 3x3 + 7 heads for L=8), leaves every
    other conv bit-identical, and restore() brings the mean-centred, clamped full-precision copy back."""
    from cunet_b200.models.cu_net import create_cu_net
    from cunet_b200.models.cu_net_prev_version import BinOp
    torch.manual_seed(0)
    net = create_cu_net(4, 32, 128, 16, 8, 1, 8, dtype="bf16")
    net.engine(1, "cuda:0")
    before = {n: p.detach().clone() for n, p in net.named_parameters()}
    op = BinOp(net)
    assert op.num_of_params == 79
    assert sorted(tuple(w.shape) for w in op.target_modules) == sorted([(32, 128, 3, 3)] * 72 + [(16, 128, 1, 1)] * 7)
    op.binarization()
    changed = [n for n, p in net.named_parameters() if not torch.equal(p.detach(), before[n])]
    assert len(changed) == 79 and all(n.endswith("conv2.weight") or n.startswith("linears.") for n in changed)
    assert "linears.7.conv.weight" not in changed
    w = dict(net.named_parameters())["hg.down_blocks.0.layers.0.conv2.weight"].detach()
    assert (w.abs() - w.abs().mean(dim=(1, 2, 3), keepdim=True)).abs().max() < 1e-6      # sign(w) * mean|w| per filter


def test_binop_matches_oracle_on_cunet_shapes():
    from cunet_b200.models.cu_net_prev_version import BinOp
    gen = torch.Generator().manual_seed(5)
    shapes = [(16, 3, 3, 3), (128, 160, 1, 1), (32, 128, 3, 3), (128, 320, 1, 1), (32, 128, 3, 3), (68, 128, 1, 1),
              (8, 8, 1, 1)]
    ws = [(torch.rand(s, generator=gen) * 2 - 1) * 1.4 for s in shapes]
    model = _model(ws)
    op = BinOp(model)
    idx = quantize_oracle.target_indices(len(ws))
    op.binarization()
    convs = [m for m in model.modules() if isinstance(m, nn.Conv2d)]
    wb, saved = quantize_oracle.binop_binarization([ws[i] for i in idx])
    for j, i in enumerate(idx):
        got = convs[i].weight.data.cpu()
        assert (got - wb[j]).abs().max() <= 1e-6 * wb[j].abs().max()
        assert torch.equal(got.sign(), wb[j].sign())
    assert torch.equal(convs[0].weight.data.cpu(), ws[0]) and torch.equal(convs[-1].weight.data.cpu(), ws[-1])
    grads = [torch.randn(s, generator=gen) * 0.01 for s in shapes]
    for m, g in zip(convs, grads):
        m.weight.grad = g.clone().cuda()
    op.restore()
    for j, i in enumerate(idx):
        assert (convs[i].weight.data.cpu() - saved[j]).abs().max() <= 1e-7
    op.updateBinaryGradWeight()
    gu = quantize_oracle.binop_update_grad(saved, [grads[i] for i in idx])
    for j, i in enumerate(idx):
        got = convs[i].weight.grad.cpu()
        assert (got - gu[j]).abs().max() <= 2e-5 * gu[j].abs().max(), i
    assert torch.equal(convs[0].weight.grad.cpu(), grads[0])


def test_quan_input_and_get_preds():
    from cunet_b200.utils.quantize import QuanInput2d
    from cunet_b200.pylib.Evaluation import get_preds
    gen = torch.Generator().manual_seed(9)
    x = (torch.randn(2, 8, 16, 16, generator=gen) * 0.8).cuda().requires_grad_(True)
    y = QuanInput2d(8)(x)
    assert torch.equal(y.detach().cpu(), quantize_oracle.quan_input_forward(x.detach().cpu(), 8))
    gy = torch.randn(2, 8, 16, 16, generator=gen).cuda()
    y.backward(gy)
    assert torch.equal(x.grad.cpu(), quantize_oracle.quan_input_backward(x.detach().cpu(), gy.cpu()))
    _, hm = synthetic.make_inputs(3, 68, seed=4)
    s = hm.clone()
    s[0, 0] = -1.0                      # masked (max <= 0)
    s[1, 1, 10, 20] = 5.0
    s[1, 1, 40, 7] = 5.0                # tie -> first maximum
    assert torch.equal(get_preds(s.cuda()).cpu(), evaluation_oracle.get_preds(s))
