"""N>1 host logic on CPU (gloo, world_size 2): the flat-bucket allreduce with 1/world pre-scaling reproduces the
single-process emulation of the reference's DataParallel step (oracle on each shard with LOCAL BatchNorm
statistics, gradients averaged) -- SURVEY.md section 4, test tier 6."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cunet_oracle
from cunet_b200 import parallel

CFG = dict(class_num=3, layer_num=2, order=1, loss_num=2, neck_size=2, growth_rate=8, init_chan_num=16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_grads(net):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).flatten() for p in net.parameters()])


def _shard_grad(state, img, hm, scale):
    # float64: this network's fp32 gradients are noisy at the 1e-3..1e-2 level (reduction-order dependent), which
    # would mask what is being tested here -- the collective plumbing
    net = cunet_oracle.OracleCUNet({k: v.double() if v.is_floating_point() else v for k, v in state.items()}, **CFG)
    for n_ in net.param_names:
        net.state[n_] = net.state[n_].detach().double().requires_grad_(True)
    loss = cunet_oracle.multi_loss_mse(net(img.double()), hm.double()) * scale   # dLoss pre-scaled by 1/world
    loss.backward()
    return _flat_grads(net)


def _worker(rank, world, port, state, img, hm, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    r, w, _ = parallel.env_world()
    start, size = parallel.shard_batch(img.shape[0], r, w)
    flat = parallel.broadcast_params(torch.cat([v.flatten() for k, v in state.items() if v.is_floating_point()]), w)
    assert flat.numel() > 0
    g = _shard_grad(state, img[start:start + size], hm[start:start + size], 1.0 / w)
    parallel.allreduce_mean(g, w)
    if rank == 0:
        torch.save(g, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_allreduce_equals_single_process_emulation(tmp_path):
    torch.manual_seed(0)
    state = cunet_oracle.init_state(CFG["class_num"], CFG["layer_num"], CFG["order"], seed=2, neck_size=2,
                                    growth_rate=8, init_chan_num=16)
    gen = torch.Generator().manual_seed(1)
    img = torch.rand(4, 3, 64, 64, generator=gen)
    hm = torch.rand(4, 3, 16, 16, generator=gen)
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), state, img, hm, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process emulation: each shard with its own batch statistics, gradients averaged
    ref = 0.5 * (_shard_grad(state, img[:2], hm[:2], 1.0) + _shard_grad(state, img[2:], hm[2:], 1.0))
    assert torch.allclose(got, ref, rtol=1e-9, atol=1e-12)
    # and it is NOT the full-batch gradient (BatchNorm statistics are per replica, like nn.DataParallel)
    full = _shard_grad(state, img, hm, 1.0)
    assert (got - full).abs().max() > 1e-6


def test_shard_batch_semantics():
    assert parallel.shard_batch(24, 3, 8) == (9, 3)          # reference: --bs is the global batch
    assert parallel.shard_batch(24, 3, 8, weak=True) == (72, 24)
    with pytest.raises(ValueError):
        parallel.shard_batch(24, 0, 5)


def _worker_sharded(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(7)
    params = torch.randn(n, generator=g, dtype=torch.float64)          # identical replicas
    sq = torch.rand(n, generator=g, dtype=torch.float64)
    grads = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank), dtype=torch.float64) / world
    lr = 2.5e-4

    # (a) allreduce + replicated RMSprop (the reference's DataParallel semantics)
    ga = parallel.allreduce_mean(grads.clone(), world)
    pa, sa = params.clone(), sq.clone()
    cunet_oracle.rmsprop_step([pa], [ga], [sa], lr)

    # (b) reduce-scatter -> RMSprop on this rank's slice only -> all-gather of the updated parameters
    s, e = parallel.shard_range(n, rank, world)
    gs = torch.empty(e - s, dtype=torch.float64)
    parallel.reduce_scatter_mean(grads.clone(), gs, world)
    assert torch.equal(gs, ga[s:e])                                     # the slice of the summed bucket, exactly
    pb, sb = params.clone(), sq.clone()
    shard_p, shard_s = pb[s:e].clone(), sb[s:e].clone()
    cunet_oracle.rmsprop_step([shard_p], [gs], [shard_s], lr)
    parallel.all_gather_params(pb, shard_p, world)
    ok = torch.equal(pb, pa) and torch.equal(shard_s, sa[s:e])
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save(dict(ok=bool(flag.item()), p=pb), out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_optimizer_exchange_equals_allreduce(tmp_path):
    """SURVEY.md section 8(f)1 host logic: reduce-scatter -> shard update -> all-gather leaves every rank with
    bit-identical parameters to allreduce + replicated RMSprop (elementwise optimizer, same summed gradients)."""
    out = str(tmp_path / "s.pt")
    mp.spawn(_worker_sharded, args=(2, _free_port(), 4096, out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
    with pytest.raises(ValueError):
        parallel.shard_range(4097, 0, 2)
    assert parallel.shard_range(4096, 1, 2) == (2048, 4096)
