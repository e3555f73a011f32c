#!/usr/bin/env python
"""Train / evaluate a CU-Net on the B200-native path -- drop-in for the reference's ``cu-net.py``.

Same flags as the reference (options/base_options.py, options/train_options.py): --layer_num --order --class_num
--loss_num --lr --bs --nEpochs --gpu_id --is_train --exp_dir --exp_id --resume_prefix --bits_w --bits_g ...
Deliberate deviations (documented in DESIGN.md):
  * boolean flags parse "false"/"0"/"no" as False (the reference's ``type=bool`` makes ``--is_train false`` True);
  * --loss_num defaults to --layer_num when the reference default (16) would violate loss_num <= layer_num;
  * multi-GPU is one process per GPU (torchrun) with one NCCL allreduce per step instead of nn.DataParallel
    (cu-net.py:59); --gpu_id selects the device of a single process;
  * the datasets (MPII / FACE image folders) are not shipped: --data synthetic (default) draws the seeded
    synthetic batches of SURVEY.md section 8(d); a torch DataLoader yielding (img, heatmap) can be plugged in
    through run(..., loader=...).
The step itself is train() / validate() of cu-net.py:147-278: net(img), multi-loss MSE, backward, RMSprop, with the
per-iteration .cpu() metric loops replaced by the fused on-device landmark decode (--fused, default) or, with
--no-fused, literally the reference's sequence through the module API (loss.backward(); optimizer.step()).
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def str2bool(v):
    return str(v).lower() not in ("false", "0", "no", "off", "")


def parse(argv=None):
    ap = argparse.ArgumentParser()
    # options/base_options.py:13-35
    ap.add_argument("--data_dir", type=str, default="./dataset")
    ap.add_argument("--exp_dir", type=str, default="./exp")
    ap.add_argument("--exp_id", type=str, default="")
    ap.add_argument("--gpu_id", type=str, default="0")
    ap.add_argument("--nThreads", type=int, default=4)
    ap.add_argument("--is_train", type=str2bool, default=True)
    ap.add_argument("--dataset", type=str, default="mpii")
    # options/train_options.py:7-38
    ap.add_argument("--layer_num", type=int, default=2)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--class_num", type=int, default=16)
    ap.add_argument("--loss_num", type=int, default=16)
    ap.add_argument("--lr", type=float, default=2.5e-4)
    ap.add_argument("--bs", type=int, default=24)
    ap.add_argument("--adjust_lr", type=str2bool, default=True,
                    help="the reference's schedule (utils/util.py:106-119, applied every epoch at cu-net.py:119)")
    ap.add_argument("--resume_prefix", type=str, default="")
    ap.add_argument("--nEpochs", type=int, default=200)
    ap.add_argument("--print_freq", type=int, default=10)
    ap.add_argument("--bits_w", type=int, default=1)
    ap.add_argument("--bits_i", type=int, default=8)
    ap.add_argument("--bits_g", type=int, default=8)
    # B200 path
    ap.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16")
    ap.add_argument("--quant", choices=["none", "bin", "quan"], default="none",
                    help="bin: BinOp protocol (cu-net-prev-version-bin.py); quan: QuanOp (…-wig.py)")
    ap.add_argument("--data", choices=["synthetic"], default="synthetic")
    ap.add_argument("--iters_per_epoch", type=int, default=20)
    ap.add_argument("--val_iters", type=int, default=2,
                    help="synthetic validation batches per epoch (validate() runs every epoch, cu-net.py:123-126); 0: skip")
    ap.add_argument("--flip_test", type=str2bool, default=True,
                    help="validation with the reference's flip test-time augmentation and heat-map PCK (cu-net.py:240-254)")
    ap.add_argument("--save_freq", type=int, default=0,
                    help="write <exp_dir>/<exp_id>/lr-<lr>-<epoch>.pth.tar (the reference's checkpoint format) every n epochs")
    ap.add_argument("--fused", dest="fused", action="store_true", default=True)
    ap.add_argument("--no-fused", dest="fused", action="store_false")
    opt = ap.parse_args(argv)
    if opt.loss_num > opt.layer_num:
        opt.loss_num = opt.layer_num
    return opt


def _package():
    from cunet_b200.utils import util
    return util


def adjust_lr(opt, epoch, optimizer=None):
    """utils/util.py:106-119 (drop-in: cunet_b200/utils/util.py)."""
    return _package().adjust_lr(opt, epoch, optimizer)


def TrainHistory():
    """utils/util.py:8-46 (drop-in: cunet_b200/utils/util.py) -- the reference's full key set."""
    return _package().TrainHistory()


def resume(net, opt, history):
    """--resume_prefix: load <exp_dir>/<exp_id>/<resume_prefix>[.pth.tar] (written by this script or by the reference,
    cu-net.py:63-73) into the network by name.  Returns the checkpoint's optimizer state (or None)."""
    import torch
    from cunet_b200.utils.checkpoint import Checkpoint
    ck = Checkpoint()
    stem = os.path.join(opt.exp_dir, opt.exp_id, opt.resume_prefix)
    ck.load_prefix = stem[:-len(".pth.tar")] if stem.endswith(".pth.tar") else stem
    if not ck.load_checkpoint(net, None, history):
        raise IOError("--resume_prefix: no checkpoint at %s.pth.tar" % ck.load_prefix)
    if history.lr:
        opt.lr = history.lr[-1]["lr"]          # the reference reads it back from the optimizer's group (cu-net.py:71)
    return torch.load(ck.load_prefix + ".pth.tar", map_location="cpu", weights_only=False).get("optimizer")


MPII_PCK_IDX = [0, 1, 2, 3, 4, 5, 10, 11, 14, 15]                      # cu-net.py:127


def validate(tr, batches, opt, dev):
    """validate() of cu-net.py:209-278 on the B200 path: eval-mode forward, multi-loss MSE of the un-flipped pass, flip
    test-time augmentation when the joint pairs are known (16 MPII joints, cu-net.py:32-33,240-248), heat-map PCK
    (:254), original-resolution PCKh and final predictions when the loader supplies the crop metadata (:256-257,275),
    everything on the device.  ``batches`` yields (img, heatmap) or the reference's 8-tuple
    (img, heatmap, center, scale, rot, grnd_pts, normalizer, index).  Returns (loss, pckh, predictions or None)."""
    import torch
    from cunet_b200.pylib import Evaluation
    from cunet_b200.utils.util import AverageMeter
    losses, pck, pck_orig = AverageMeter(), AverageMeter(), AverageMeter()
    idx = [i for i in MPII_PCK_IDX if i < opt.class_num]
    preds_all = []
    tr.net.eval()
    with torch.no_grad():
        for batch in batches:
            img, hm = batch[0].to(dev, non_blocking=True), batch[1].to(dev, non_blocking=True)
            if opt.class_num == 16 and opt.flip_test:
                loss, _, out = tr.eval_step_flip(img, hm)
            else:
                loss, _ = tr.eval_step(img, hm)
                out = tr.eng.head_outputs()[-1].float().contiguous()
            losses.update(float(loss))
            pck.update(float(Evaluation.accuracy(out, hm, idx)[0]))
            if len(batch) >= 8:
                center, scale, rot, grnd, norm, index = batch[2:8]
                pck_orig.update(float(Evaluation.accuracy_origin_res(out, center, scale, [64, 64], grnd, norm, rot)[0]))
                preds_all.append((index, Evaluation.final_preds(out, center, scale, [64, 64], rot).cpu()))
    tr.net.train()
    predictions = None
    if preds_all:
        n = 1 + max(int(i.max()) for i, _ in preds_all)
        predictions = torch.zeros(n, opt.class_num, 2)
        for index, p in preds_all:
            predictions[index.long()] = p
    return losses.avg, (pck_orig.avg if pck_orig.count else pck.avg), predictions


def run(opt, loader=None, val_loader=None):
    import torch
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):
        ge.build()
    from cunet_b200.models.cu_net import create_cu_net
    from cunet_b200.engine import Trainer
    from cunet_b200.utils.quantize import BinOp, QuanOp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", opt.gpu_id.split(",")[0]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    torch.manual_seed(0)
    # --quant quan = cu-net-prev-version-wig.py: QuanOp on the weights / gradients AND the activation-quantized model
    # (QuanInput2d with --bits_i bits in front of the 3x3 and head convs, models/cu_net_prev_version_wig.py:96-98,277-279)
    bits_i = opt.bits_i if (opt.quant == "quan" and opt.bits_i <= 8) else 0
    net = create_cu_net(neck_size=4, growth_rate=32, init_chan_num=128, class_num=opt.class_num,
                        layer_num=opt.layer_num, order=opt.order, loss_num=opt.loss_num, dtype=opt.dtype, bits_i=bits_i)
    history = TrainHistory()
    opt_state = resume(net, opt, history) if opt.resume_prefix else None     # host-side, before the weights move to HBM
    from cunet_b200.parallel import shard_batch
    _, bs = shard_batch(opt.bs, rank, world)  # the reference's DataParallel splits --bs over the GPUs (cu-net.py:59,84)
    if not opt.fused and (opt.quant != "none" or world > 1):
        raise SystemExit("--no-fused (the literal module-API step of cu-net.py:171-183) supports neither --quant nor "
                         "WORLD_SIZE > 1: the quantizer protocol and the gradient allreduce live in the fused Trainer")
    eng = net.engine(bs, dev)
    quant = None
    if opt.quant == "bin":
        quant = BinOp(net)
    elif opt.quant == "quan":
        quant = QuanOp(net, bits_w=opt.bits_w, bits_g=opt.bits_g)
    tr = Trainer(net, bs, lr=opt.lr, device=dev, process_group=pg, world_size=world, quant=quant,
                 use_graph=opt.fused)      # the benchmarked path: CUDA-graph replay of the fused step
    if opt_state is not None and opt_state.get("state") and opt.fused:
        from cunet_b200.utils.checkpoint import rmsprop_state_to_flat
        rmsprop_state_to_flat(net, opt_state, eng)
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(eng.params, 0)
    if loader is None:
        from cunet_b200.utils import synthetic   # seeded synthetic batches

        def loader_fn(epoch):
            for it in range(opt.iters_per_epoch):
                yield synthetic.make_inputs(bs, opt.class_num, seed=1000 * epoch + it * world + rank)

        def val_loader_fn():
            for it in range(opt.val_iters):
                yield synthetic.make_inputs(bs, opt.class_num, seed=900000 + it * world + rank)
    else:
        def loader_fn(epoch):
            return iter(loader)

        def val_loader_fn():
            return iter(val_loader if val_loader is not None else [])
    losses = []
    if not opt.is_train:
        # cu-net.py:101-107: one validation pass (quantized weights when a quantizer is configured)
        if quant is not None:
            (quant.binarization if opt.quant == "bin" else quant.quantization)()
        val_loss, val_pckh, predictions = validate(tr, val_loader_fn() if val_loader is not None or opt.val_iters
                                                   else loader_fn(0), opt, dev)
        if quant is not None:
            quant.restore()
        if rank == 0:
            print("val loss %.6f, pck %.4f" % (val_loss, val_pckh))
        return [val_loss]
    net.train()
    opt_torch = None
    if not opt.fused:
        opt_torch = torch.optim.RMSprop(net.parameters(), lr=opt.lr, alpha=0.99, eps=1e-8, momentum=0, weight_decay=0)
        if opt_state is not None and opt_state.get("state"):
            opt_torch.load_state_dict(opt_state)       # cu-net.py:70
    from collections import OrderedDict
    for epoch in range(history.last_epoch() + 1, opt.nEpochs):
        if opt.adjust_lr:
            tr.set_lr(adjust_lr(opt, epoch, opt_torch))
        t0, n_img, last, tot, cnt = time.time(), 0, 0.0, 0.0, 0
        for i, (img, hm) in enumerate(loader_fn(epoch)):
            if opt.fused:
                last = float(tr.train_step(img.to(dev, non_blocking=True), hm.to(dev, non_blocking=True)))
            else:   # the reference's literal step through the module API (cu-net.py:171-183)
                out = net(img.to(dev))
                loss = 0
                for o in out:
                    d = (o - hm.to(dev)) ** 2
                    loss = loss + d.sum() / d.numel()
                opt_torch.zero_grad()
                loss.backward()
                opt_torch.step()
                last = float(loss.detach())
            n_img += img.shape[0] * world
            tot, cnt = tot + last, cnt + 1
            if rank == 0 and (i % opt.print_freq == 0):
                print("epoch %d iter %d loss %.6f" % (epoch, i, last))
        losses.append(last)
        dt = time.time() - t0
        # cu-net.py:123-139: validate every epoch, record the history, keep the best checkpoint
        val_loss, val_pckh, predictions = validate(tr, val_loader_fn(), opt, dev) if opt.val_iters or val_loader is not None \
            else (0.0, 0.0, None)
        history.update(OrderedDict([("epoch", epoch)]), OrderedDict([("lr", opt.lr)]),
                       OrderedDict([("train_loss", tot / max(1, cnt)), ("val_loss", val_loss)]),
                       OrderedDict([("val_pckh", val_pckh)]))
        if rank == 0:
            print("epoch %d done: train loss %.6f, val loss %.6f, val pck %.4f%s, %.1f images/s" %
                  (epoch, tot / max(1, cnt), val_loss, val_pckh, " (best)" if history.is_best else "", n_img / dt))
            if opt.save_freq > 0 and (epoch + 1) % opt.save_freq == 0:
                from cunet_b200.utils.checkpoint import Checkpoint, rmsprop_state_from_flat
                ck = Checkpoint()
                os.makedirs(os.path.join(opt.exp_dir, opt.exp_id), exist_ok=True)
                ck.save_prefix = os.path.join(opt.exp_dir, opt.exp_id) + os.sep
                state = opt_torch.state_dict() if opt_torch is not None else rmsprop_state_from_flat(net, eng, opt.lr)
                ck.save_checkpoint(net, state, history, predictions)
    return losses


if __name__ == "__main__":
    run(parse())
