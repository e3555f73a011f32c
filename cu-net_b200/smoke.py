"""One tiny invocation of the hot path on cuda:0 checked against the CPU oracle (driver smoke test)."""
import os
import sys

import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from cunet_b200.models.cu_net import create_cu_net
    from cunet_b200.engine import Trainer
    from oracle import cunet_oracle, evaluation_oracle, synthetic          # checker only

    class_num, L, K, loss_num, n = 16, 2, 1, 2, 1
    torch.manual_seed(0)
    net = create_cu_net(4, 32, 128, class_num, L, K, loss_num, dtype="fp32")
    img, hm = synthetic.make_inputs(n, class_num, seed=0)
    tr = Trainer(net, n, lr=2.5e-4, device="cuda:0")
    state = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    loss = tr.train_step(img.cuda(), hm.cuda())
    torch.cuda.synchronize()
    ora = cunet_oracle.OracleCUNet(state, class_num, L, K, loss_num)
    outs = ora(img)
    oloss = cunet_oracle.multi_loss_mse(outs, hm)
    got = [o.cpu() for o in tr.eng.head_outputs()]
    rel = max(((g - o.detach()).abs().max() / o.detach().abs().max()).item() for g, o in zip(got, outs))
    lrel = abs(float(loss) - float(oloss)) / abs(float(oloss))
    preds = tr.eng.preds.cpu()
    opreds = evaluation_oracle.get_preds(got[-1])
    print("smoke: loss %.6f (oracle %.6f, rel %.2e)  head rel err %.2e  decode exact %s" %
          (float(loss), float(oloss), lrel, rel, bool(torch.equal(preds, opreds))))
    assert rel < 5e-3 and lrel < 5e-3, "CUDA path disagrees with the oracle"
    assert torch.equal(preds, opreds), "argmax decode mismatch"
