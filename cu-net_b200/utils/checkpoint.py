"""Drop-in for the reference's ``utils/checkpoint.py`` (Checkpoint.save_checkpoint :13-31, load_checkpoint :40-67).

Same file format: ``torch.save({'train_history': ..., 'state_dict': net.state_dict(), 'optimizer': ...})`` under
``<save_prefix>lr-<lr>-<epoch>.pth.tar``, weights copied back BY NAME.  Two things differ from the reference's
in-place loop, both needed to load the checkpoints it wrote itself:

* reference checkpoints are saved from an ``nn.DataParallel`` wrapper, so every key carries a ``module.`` prefix
  (cu-net.py:59; the commented ``name = name[7:]`` at utils/checkpoint.py:55 is exactly this); ``CUNetB200`` is not
  wrapped (one process per GPU), so the prefix is stripped on load and, on request, added on save;
* PyTorch 0.4-era checkpoints have no ``num_batches_tracked`` buffers: missing keys keep their current values.

The tensors are NCHW fp32 exactly as the reference stores them; the tensor-core operand images are derived from
them by ``cunet_pack_weights`` at the start of the next step, so loading is a plain copy into the (host or device)
parameter storage.  The optimizer entry is ``torch.optim.RMSprop.state_dict()``; for the fused ``Trainer`` (flat
``square_avg`` buffer) use ``rmsprop_state_to_flat`` / ``rmsprop_state_from_flat``.
"""
import os
import shutil

import torch

PREFIX = "module."


def strip_prefix(state_dict):
    """{'module.x': t} -> {'x': t} (keys without the prefix pass through)."""
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state_dict.items()}


def load_weights(net, state_dict, strict=False):
    """Copy ``state_dict`` into ``net`` by name (utils/checkpoint.py:53-61).  Returns (loaded, skipped, missing):
    names copied, names in the file the net does not have, names of the net the file does not have."""
    src = strip_prefix(state_dict)
    dst = net.state_dict()
    loaded, skipped = [], []
    with torch.no_grad():
        for name, value in src.items():
            if name not in dst:
                skipped.append(name)
                continue
            value = value.data if isinstance(value, torch.nn.Parameter) else value
            if tuple(dst[name].shape) != tuple(value.shape):
                raise ValueError("checkpoint tensor %s has shape %s, the network expects %s" %
                                 (name, tuple(value.shape), tuple(dst[name].shape)))
            dst[name].copy_(value)
            loaded.append(name)
    missing = [k for k in dst if k not in src]
    if strict and (skipped or [m for m in missing if not m.endswith("num_batches_tracked")]):
        raise KeyError("checkpoint / network mismatch: not in net %s, not in file %s" % (skipped[:5], missing[:5]))
    return loaded, skipped, missing


def rmsprop_state_to_flat(net, optimizer_state, eng):
    """torch.optim.RMSprop.state_dict() -> the engine's flat ``sq_avg`` buffer (parameter order = net.parameters())."""
    names = [n for n, _ in net.named_parameters()]
    for idx, name in enumerate(names):
        st = optimizer_state["state"].get(idx)
        if st is None:
            continue
        o, n, shape = eng.p_off[name]
        eng.sq_avg[o:o + n].view(shape).copy_(st["square_avg"])


def rmsprop_state_from_flat(net, eng, lr, alpha=0.99, eps=1e-8, step=0):
    """The engine's flat ``sq_avg`` -> a state_dict torch.optim.RMSprop.load_state_dict accepts."""
    names = [n for n, _ in net.named_parameters()]
    state = {}
    for idx, name in enumerate(names):
        o, n, shape = eng.p_off[name]
        state[idx] = {"step": step, "square_avg": eng.sq_avg[o:o + n].view(shape).detach().cpu().clone()}
    group = dict(lr=lr, momentum=0, alpha=alpha, eps=eps, centered=False, weight_decay=0, params=list(range(len(names))))
    return {"state": state, "param_groups": [group]}


class Checkpoint(object):
    def __init__(self, data_parallel_names=True):
        """data_parallel_names: write ``module.``-prefixed keys, like the reference's DataParallel-wrapped net does,
        so the files load in the reference unchanged."""
        self.save_prefix = ""
        self.load_prefix = ""
        self.data_parallel_names = data_parallel_names

    def _paths(self, train_history):
        lr_prefix = ("lr-%.15f" % train_history.lr[-1]["lr"]).rstrip("0").rstrip(".")
        epoch = train_history.epoch[-1]["epoch"]
        return self.save_prefix + lr_prefix + ("-%d" % epoch)

    def save_checkpoint(self, net, optimizer, train_history, preds=None):
        """utils/checkpoint.py:13-31 (the predictions go to a .mat file next to it, as in the reference)."""
        stem = self._paths(train_history)
        state = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        if self.data_parallel_names:
            state = {PREFIX + k: v for k, v in state.items()}
        opt_state = optimizer.state_dict() if hasattr(optimizer, "state_dict") else optimizer
        torch.save({"train_history": train_history.state_dict(), "state_dict": state, "optimizer": opt_state},
                   stem + ".pth.tar")
        if preds is not None:
            import scipy.io
            scipy.io.savemat(stem + "-preds.mat", mdict={"preds": preds.detach().cpu().numpy()})
        if getattr(train_history, "is_best", False):
            shutil.copyfile(stem + ".pth.tar", stem + "-model-best.pth.tar")
            if preds is not None:
                shutil.copyfile(stem + "-preds.mat", stem + "-preds-best.mat")
        return stem + ".pth.tar"

    def load_checkpoint(self, net, optimizer, train_history):
        """utils/checkpoint.py:40-67.  Returns True when a checkpoint was found."""
        path = self.load_prefix + ".pth.tar"
        if not os.path.isfile(path):
            print("=> no checkpoint found at '{}'".format(path))
            return False
        print("=> loading checkpoint '{}'".format(path))
        ck = torch.load(path, map_location="cpu", weights_only=False)
        if train_history is not None and "train_history" in ck:
            train_history.load_state_dict(ck["train_history"])
        if optimizer is not None and "optimizer" in ck and hasattr(optimizer, "load_state_dict"):
            optimizer.load_state_dict(ck["optimizer"])
        _, skipped, _ = load_weights(net, ck["state_dict"])
        for name in skipped:
            print("=> not load weights '{}'".format(name))
        return True
