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
