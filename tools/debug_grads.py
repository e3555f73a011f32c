import sys, torch
sys.path.insert(0, '.')
from tests.test_gpu_model import _setup, _rel
from oracle import cunet_oracle
cfg = tuple(int(x) for x in sys.argv[1:6]) if len(sys.argv) > 5 else (68, 2, 1, 2, 2)
dtype = sys.argv[6] if len(sys.argv) > 6 else "fp32"
class_num, L, K, loss_num, n = cfg
net, ora, img, hm = _setup(class_num, L, K, loss_num, n, dtype)
net.train()
outs = net(img.cuda())
loss = cunet_oracle.multi_loss_mse(outs, hm.cuda())
loss.backward()
oouts = ora(img)
oloss = cunet_oracle.multi_loss_mse(oouts, hm)
oloss.backward()
# exact (fp64) oracle
state = cunet_oracle.init_state(class_num, L, K, seed=0)
o64 = cunet_oracle.OracleCUNet({k: v.double() if v.is_floating_point() else v for k, v in state.items()}, class_num, L, K, loss_num)
for nme in o64.param_names: o64.state[nme] = o64.state[nme].detach().double().requires_grad_(True)
outs64 = o64(img.double()); l64 = cunet_oracle.multi_loss_mse(outs64, hm.double()); l64.backward()
def rms(a, b):
    a, b = a.double(), b.double()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
for a, b, c in zip(outs, oouts, outs64):
    print("head rms-rel mine-vs-64 %.3e" % rms(a.detach().cpu(), c.detach()))
    print("head: mine-vs-64 %.3e   oracle32-vs-64 %.3e" % (_rel(a.detach().cpu(), c.detach()), _rel(b.detach(), c.detach())))
print("loss mine %.9f o32 %.9f o64 %.9f" % (float(loss.detach()), float(oloss.detach()), float(l64.detach())))
mine, o32 = [], []
for name, p in net.named_parameters():
    g64 = o64.state[name].grad
    if g64 is None: continue
    mine.append((_rel(p.grad.cpu(), g64), name)); o32.append((_rel(ora.state[name].grad, g64), name))
ms = sorted(m[0] for m in mine); os_ = sorted(o[0] for o in o32)
print("grad err vs fp64:  mine median %.2e p90 %.2e max %.2e | oracle32 median %.2e p90 %.2e max %.2e" % (
    ms[len(ms)//2], ms[int(len(ms)*0.9)], ms[-1], os_[len(os_)//2], os_[int(len(os_)*0.9)], os_[-1]))
cos = sorted(torch.nn.functional.cosine_similarity(p.grad.cpu().flatten().double(), o64.state[n_].grad.flatten(), dim=0).item() for n_, p in net.named_parameters() if o64.state[n_].grad is not None and p.numel() >= 64)
print("grad cosine vs fp64: min %.4f p10 %.4f median %.4f" % (cos[0], cos[len(cos)//10], cos[len(cos)//2]))
print("worst mine:", sorted(mine, reverse=True)[:5])
print("worst o32 :", sorted(o32, reverse=True)[:5])
