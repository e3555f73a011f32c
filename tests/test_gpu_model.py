"""Whole-model GPU parity against the CPU oracle through the reference-facing API (create_cu_net)."""
import pytest
import torch

from oracle import cunet_oracle, evaluation_oracle, synthetic

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20)).item()


def _setup(class_num, L, K, loss_num, n, dtype, seed=0):
    from cunet_b200.models.cu_net import create_cu_net
    torch.manual_seed(seed)
    net = create_cu_net(4, 32, 128, class_num, L, K, loss_num, dtype=dtype)
    state = cunet_oracle.init_state(class_num, L, K, seed=seed)
    img, hm = synthetic.make_inputs(n, class_num, seed=seed)
    net.engine(n, "cuda:0")                       # binds parameters to device storage
    net.load_state_dict(state)
    ora = cunet_oracle.OracleCUNet(state, class_num, L, K, loss_num)
    return net, ora, img, hm


def _oracle64(class_num, L, K, loss_num, img, hm, seed=0):
    """Exact (float64) oracle gradients: the yardstick for both the CUDA path and the fp32 oracle."""
    state = cunet_oracle.init_state(class_num, L, K, seed=seed)
    o64 = cunet_oracle.OracleCUNet({k: v.double() if v.is_floating_point() else v for k, v in state.items()},
                                   class_num, L, K, loss_num)
    for nme in o64.param_names:
        o64.state[nme] = o64.state[nme].detach().double().requires_grad_(True)
    outs64 = o64(img.double())
    l64 = cunet_oracle.multi_loss_mse(outs64, hm.double())
    l64.backward()
    return o64, [o.detach() for o in outs64]


@pytest.mark.parametrize("cfg", [(68, 2, 1, 2, 2), (16, 3, 2, 2, 1), (5, 3, 0, 3, 2)], ids=["L2K1C68", "L3K2C16", "L3K0C5"])
def test_fp32_forward_backward_parity(cfg):
    """fp32 storage, 3xTF32 tensor-core arithmetic.

    Outputs and loss: within 1e-3 of the fp32 oracle (measured ~1e-4).
    Gradients: this network's gradients are ill-conditioned -- the fp32 ORACLE ITSELF deviates from the exact
    float64 gradients by ~1e-2 (median over tensors) at these batch sizes, so "1e-3 vs the fp32 oracle" is not a
    meaningful bar.  The CUDA path is therefore held to the exact float64 gradients with a budget of 5x the fp32
    oracle's own error (measured ~3x: its forward error is ~3x the fp32 oracle's), see DESIGN.md "Numerics"."""
    class_num, L, K, loss_num, n = cfg
    net, ora, img, hm = _setup(class_num, L, K, loss_num, n, "fp32")
    net.train()
    outs = net(img.cuda())
    loss = cunet_oracle.multi_loss_mse(outs, hm.cuda())
    loss.backward()
    oouts = ora(img)
    oloss = cunet_oracle.multi_loss_mse(oouts, hm)
    oloss.backward()
    assert len(outs) == loss_num
    o64, outs64 = _oracle64(class_num, L, K, loss_num, img, hm)
    for a, b, c in zip(outs, oouts, outs64):
        assert tuple(a.shape) == tuple(b.shape)
        if L == 2:      # the parity configuration (BASELINE.json configs[1]): 1e-3 against the fp32 oracle
            assert _rel(a.detach().cpu(), b.detach()) < 1e-3
        # deeper stacks amplify fp32 rounding ~10-40x per U-Net in the oracle itself: hold the CUDA path to the
        # exact float64 result with a budget of 8x the fp32 oracle's own deviation
        assert _rel(a.detach().cpu(), c) < max(1e-3, 8 * _rel(b.detach(), c))
    assert abs(float(loss.detach()) - float(oloss.detach())) / abs(float(oloss.detach())) < 1e-3
    mine, o32 = [], []
    for name, p in net.named_parameters():
        g64 = o64.state[name].grad
        if g64 is None:
            assert p.grad is None or p.grad.abs().max() == 0, name
            continue
        mine.append(_rel(p.grad.cpu(), g64))
        o32.append(_rel(ora.state[name].grad, g64))
        if p.numel() >= 64:
            cos = torch.nn.functional.cosine_similarity(p.grad.cpu().flatten().double(), g64.flatten(), dim=0)
            assert cos > 0.97, (name, float(cos))
    mine.sort()
    o32.sort()
    med, p90 = mine[len(mine) // 2], mine[int(len(mine) * 0.9)]
    omed, op90 = o32[len(o32) // 2], o32[int(len(o32) * 0.9)]
    assert med < max(2e-3, 5 * omed), "median grad err %g (fp32 oracle's own: %g)" % (med, omed)
    assert p90 < max(5e-3, 5 * op90), "p90 grad err %g (fp32 oracle's own: %g)" % (p90, op90)
    sd = net.state_dict()
    for k, v in ora.state.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert _rel(sd[k].cpu(), v) < 2e-3, k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k


def test_eval_mode_and_decode_exact():
    net, ora, img, hm = _setup(68, 2, 1, 2, 2, "fp32")
    net.eval()
    ora.eval()
    with torch.no_grad():
        outs = net(img.cuda())
        oouts = ora(img)
    for a, b in zip(outs, oouts):
        assert _rel(a.cpu(), b) < 1e-3
    from cunet_b200.engine import Trainer
    tr = Trainer(net, 2, device="cuda:0")
    loss, preds = tr.eval_step(img.cuda(), hm.cuda())
    got_last = tr.eng.head_outputs()[-1].cpu()
    # decode kernel vs the reference's get_preds on the SAME heatmaps: bit exact
    assert torch.equal(preds.cpu(), evaluation_oracle.get_preds(got_last))
    # and on the synthetic targets (unique, well separated peaks): equal to the oracle end to end
    tr.eng.act[tr.eng.plan.heads[-1].name].view(2, 64, 64, -1)[..., :68].copy_(hm.cuda().permute(0, 2, 3, 1))
    tr.eng.loss_and_decode(with_grad=False)
    assert torch.equal(tr.eng.preds.cpu(), evaluation_oracle.get_preds(hm))
    oloss = cunet_oracle.multi_loss_mse(oouts, hm)
    assert abs(float(loss) - float(oloss.detach())) / abs(float(oloss.detach())) < 1e-3


def test_fused_train_step_loss_and_rmsprop():
    """Trainer.train_step: fused loss == oracle loss; fused RMSprop == torch.optim.RMSprop's formula
    (cu-net.py:60-61) applied to the gradients the backward kernels produced."""
    class_num, L, K, loss_num, n = 16, 2, 1, 2, 2
    net, ora, img, hm = _setup(class_num, L, K, loss_num, n, "fp32")
    from cunet_b200.engine import Trainer
    tr = Trainer(net, n, lr=2.5e-4, device="cuda:0")
    e = tr.eng
    tr.load_batch(img.cuda(), hm.cuda())
    e.forward(train=True)
    e.loss_and_decode(with_grad=True)
    e.backward()
    torch.cuda.synchronize()
    g = e.grads.clone()
    p0 = e.params.clone()
    e.optimizer_step()
    torch.cuda.synchronize()
    v = 0.01 * g * g
    expect = p0 - 2.5e-4 * g / (v.sqrt() + 1e-8)
    # fp32 rounding only: the kernel forms (1-alpha) in fp32 and fuses the multiply-adds
    assert _rel(e.params, expect) < 1e-5
    assert _rel(e.sq_avg, v) < 1e-5
    oloss = cunet_oracle.multi_loss_mse(ora(img), hm)
    assert abs(float(e.loss_value()) - float(oloss.detach())) / abs(float(oloss.detach())) < 1e-3
    # second step runs on the updated weights and lowers nothing catastrophically (smoke for re-packing)
    l2 = tr.train_step()
    assert torch.isfinite(l2)
    # a torch optimizer on the module's parameters sees the same storage (drop-in path of cu-net.py:60)
    opt = torch.optim.RMSprop(net.parameters(), lr=2.5e-4, alpha=0.99, eps=1e-8)
    assert sum(p.numel() for p in net.parameters()) == 1923200       # SURVEY.md section 6 (L=2, 16 classes)
    del opt


def test_bf16_forward_backward_tolerance():
    """bf16 storage / bf16 tensor-core operands.  NOT a parity claim: the network at random init amplifies a 2^-9
    perturbation of the stored activations to ~5% / ~15% rms at the two heads even inside the fp32 oracle
    (tests/test_sensitivity_cpu.py reproduces that on the CPU), so the bar is that the CUDA bf16 path stays
    within 2x of that inherent deviation and the loss within 1e-2."""
    class_num, L, K, loss_num, n = 68, 2, 1, 2, 2
    net, ora, img, hm = _setup(class_num, L, K, loss_num, n, "bf16")
    net.train()
    outs = net(img.cuda())
    loss = cunet_oracle.multi_loss_mse(outs, hm.cuda())
    loss.backward()
    oouts = ora(img)
    oloss = cunet_oracle.multi_loss_mse(oouts, hm)

    def rms(a, b):
        a, b = a.double(), b.double()
        return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
    errs = [rms(a.detach().cpu(), b.detach()) for a, b in zip(outs, oouts)]
    assert errs[0] < 0.12 and errs[1] < 0.35, errs
    assert abs(float(loss.detach()) - float(oloss.detach())) / abs(float(oloss.detach())) < 1e-2
    for name, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), name


def test_flip_tta_eval_step_matches_oracle():
    """Trainer.eval_step_flip == validate() of cu-net.py:225-249 computed with the oracle on the CPU (eval-mode BN)."""
    from cunet_b200.engine import Trainer
    from cunet_b200.pylib import HumanAug
    class_num, L, K, loss_num, n = 16, 2, 1, 2, 2
    net, ora, img, hm = _setup(class_num, L, K, loss_num, n, "fp32")
    net.eval()
    ora.eval()
    tr = Trainer(net, n, device="cuda:0")
    loss, preds, avg = tr.eval_step_flip(img.cuda(), hm.cuda())
    with torch.no_grad():
        o1 = ora(img)
        o2 = ora(torch.flip(img, dims=[3]))
        want = (o1[-1] + evaluation_oracle.shuffle_channels_for_horizontal_flipping(
            evaluation_oracle.flip_channels(o2[-1]), HumanAug.MPII_FLIP_INDEX)) / 2
        oloss = cunet_oracle.multi_loss_mse(o1, hm)
    assert _rel(avg.cpu(), want) < 1e-3
    assert abs(float(loss) - float(oloss)) < 1e-3 * abs(float(oloss))
    assert torch.equal(preds.cpu(), evaluation_oracle.get_preds(avg.cpu()))


def test_step_is_repeatable_within_accumulation_order_noise():
    """dW and the gradient accumulators are summed with floating-point atomics / L2 reduce-adds whose order varies from
    run to run: the forward pass is bit-reproducible, and two backward passes from the same state
    agree to rounding-noise level -- this pins how large that noise is allowed to be."""
    from cunet_b200.engine import Trainer
    class_num, L, K, loss_num, n = 16, 2, 1, 2, 4
    for dtype, cos_min, rel_max in (("fp32", 0.999999, 1e-3), ("bf16", 0.9995, 5e-2)):
        net, _, img, hm = _setup(class_num, L, K, loss_num, n, dtype)
        tr = Trainer(net, n, device="cuda:0")
        e = tr.eng
        tr.load_batch(img.cuda(), hm.cuda())
        runs = []
        for _ in range(2):
            e.forward(train=True)
            e.loss_and_decode(with_grad=True)
            e.backward()
            torch.cuda.synchronize()
            runs.append((float(e.loss_value()), [o.clone() for o in e.head_outputs()], e.grads.clone()))
        assert runs[0][0] == runs[1][0], dtype                                   # loss: bit identical
        for a, b in zip(runs[0][1], runs[1][1]):
            assert torch.equal(a, b), dtype                                      # heads: bit identical
        g0, g1 = runs[0][2].double(), runs[1][2].double()
        cos = torch.nn.functional.cosine_similarity(g0, g1, dim=0).item()
        rel = ((g0 - g1).norm() / g0.norm()).item()
        print("repeatability %s: cos %.8f rel %.3e" % (dtype, cos, rel))
        assert cos > cos_min and rel < rel_max, (dtype, cos, rel)


def test_two_gpu_data_parallel_step_over_nccl():
    """pytest -m gpu on a box with >= 2 GPUs: the 2-rank NCCL step equals the single-process emulation and the
    reduce-scatter / shard-optimizer / all-gather variant equals the allreduce form (tools/dp_check.py)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29671", os.path.join(root, "tools", "dp_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "DP_CHECK world=2" in out.stdout and "FAIL" not in out.stdout, out.stdout
    assert "SHARD_OPT_CHECK world=2 params identical on every rank: OK" in out.stdout, out.stdout


def test_wig_model_activation_quantization_matches_oracle():
    """The activation-quantized model of cu-net-prev-version-wig.py (QuanInput2d in front of every 3x3 and head conv,
    models/cu_net_prev_version_wig.py:96-98,277-279): fused into the convs' operand transform here, against the CPU oracle
    with the same quantizer (forward and straight-through backward).  fp32 storage; a rounding tie of the 1/128 grid may
    fall on the other side (3xTF32 vs fp32 accumulation), so the bars are a few grid steps wide."""
    from cunet_b200.models.cu_net_prev_version_wig import create_cu_net
    class_num, L, K, loss_num, n = 16, 2, 1, 2, 2
    state = cunet_oracle.init_state(class_num, L, K, seed=0)
    img, hm = synthetic.make_inputs(n, class_num, seed=0)
    net = create_cu_net(4, 32, 128, class_num, L, K, loss_num, dtype="fp32", bits_i=8)
    assert net.plan.quan_input_bits == 8
    net.engine(n, "cuda:0")
    net.load_state_dict(state)
    net.train()
    outs = net(img.cuda())
    loss = cunet_oracle.multi_loss_mse(outs, hm.cuda())
    loss.backward()
    ora = cunet_oracle.OracleCUNet(state, class_num, L, K, loss_num, quan_input_bits=8)
    oouts = ora(img)
    oloss = cunet_oracle.multi_loss_mse(oouts, hm)
    oloss.backward()
    plain = cunet_oracle.OracleCUNet(state, class_num, L, K, loss_num)(img)
    errs = [_rel(a.detach().cpu(), b.detach()) for a, b in zip(outs, oouts)]
    gaps = [_rel(c.detach(), b.detach()) for c, b in zip(plain, oouts)]
    print("wig heads rel err", errs, "quantized-vs-plain gap", gaps, "loss", float(loss), float(oloss))
    for e_, g_ in zip(errs, gaps):
        # a rounding tie that falls on the other side moves one activation by a whole grid step (1/128), and the network
        # amplifies that like any other perturbation (~3.5x per U-Net): measured 1.1 % / 6.3 % at the two heads, against a
        # 75 % / 91 % gap to the unquantized model
        assert e_ < 0.12 and e_ < 0.15 * g_, (errs, gaps)
    assert abs(float(loss.detach()) - float(oloss.detach())) < 1e-2 * abs(float(oloss.detach()))
    coss = []
    for name, p in net.named_parameters():
        g = ora.state[name].grad
        if g is not None and p.numel() >= 1024:
            coss.append(torch.nn.functional.cosine_similarity(p.grad.cpu().flatten().double(), g.flatten().double(), dim=0).item())
    coss.sort()
    print("wig grad cosine: min %.4f median %.4f" % (coss[0], coss[len(coss) // 2]))
    # measured: min 0.81, median 0.83 (a grid-step flip of a quantized activation is a large, discontinuous change, and
    # the network amplifies it; without the quantizer the same comparison gives > 0.97, test_fp32_forward_backward_parity)
    assert coss[len(coss) // 2] > 0.7 and coss[0] > 0.6, coss[:5]
