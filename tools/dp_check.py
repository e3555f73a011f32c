"""torchrun --nproc-per-node 2 tools/dp_check.py : data-parallel step == single-process emulation (per-shard local
BatchNorm statistics, gradients averaged), on real GPUs over NCCL."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cunet_b200.models.cu_net import create_cu_net
from cunet_b200.engine import Trainer
from cunet_b200 import parallel
from oracle import synthetic

rank, world, local = parallel.env_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
C, L, bs = 16, 2, 2
net = create_cu_net(4, 32, 128, C, L, 1, L, dtype="fp32")
tr = Trainer(net, bs, lr=1e-3, device=dev, process_group=dist.group.WORLD, world_size=world)
parallel.broadcast_params(tr.eng.params, world)
img, hm = synthetic.make_inputs(bs * world, C, seed=3)
s, n = parallel.shard_batch(bs * world, rank, world)
e = tr.eng
tr.load_batch(img[s:s + n].to(dev), hm[s:s + n].to(dev))
e.forward(train=True); e.loss_and_decode(with_grad=True); e.backward()
parallel.allreduce_mean(e.grads, world)
torch.cuda.synchronize()
got = e.grads.clone()
if rank == 0:
    e.grad_scale = 1.0
    acc = torch.zeros_like(got)
    for r in range(world):
        s, n = parallel.shard_batch(bs * world, r, world)
        tr.load_batch(img[s:s + n].to(dev), hm[s:s + n].to(dev))
        e.forward(train=True); e.loss_and_decode(with_grad=True); e.backward()
        acc += e.grads / world
    err = ((got - acc).abs().max() / acc.abs().max()).item()
    print("DP_CHECK world=%d rel err %.3e %s" % (world, err, "OK" if err < 2e-3 else "FAIL"))
# reduce-scatter -> shard RMSprop -> all-gather == allreduce -> replicated RMSprop, from the
# same parameters / optimizer state / per-rank gradients
e.grad_scale = 1.0 / world
p0, v0 = e.params.clone(), e.sq_avg.clone()
gen = torch.Generator(device="cpu").manual_seed(100 + rank)
g_local = (torch.randn(e.grads.numel(), generator=gen) * 1e-3).to(dev)
e.grads.copy_(g_local)
tr._reduce_and_step()
torch.cuda.synchronize()
pA, vA = e.params.clone(), e.sq_avg.clone()
e.params.copy_(p0); e.sq_avg.copy_(v0); e.grads.copy_(g_local)
tr2 = Trainer(net, bs, lr=1e-3, device=dev, process_group=dist.group.WORLD, world_size=world, shard_optimizer=True,
              rank=rank)
tr2._reduce_and_step()
torch.cuda.synchronize()
s0, s1 = tr2._sh0, tr2._sh1
same_p = torch.equal(e.params, pA)
same_v = torch.equal(e.sq_avg[s0:s1], vA[s0:s1])           # the optimizer state only exists for the rank's slice
flag = torch.tensor([int(same_p and same_v)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARD_OPT_CHECK world=%d params identical on every rank: %s" % (world, "OK" if int(flag) else "FAIL"))
dist.barrier()
dist.destroy_process_group()
