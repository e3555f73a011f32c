"""Drop-in for the reference's ``models/cu_net.py``: ``create_cu_net(neck_size, growth_rate, init_chan_num,
class_num, layer_num, order, loss_num) -> nn.Module`` (models/cu_net.py:362-368).

Same constructor contract (loss anchors :274-283, ``order < layer_num`` :285-287), same ``state_dict`` names
and NCHW shapes (so reference checkpoints load, utils/checkpoint.py:40-67), same ``forward(x) -> list`` of
``loss_num`` heatmaps [N, class_num, 64, 64], ``.train()/.eval()``, ``.parameters()`` -- but the arithmetic runs
in the sm_100a kernels of libcunet_b200.so through ``engine.Engine``.  The module tree exists only to hold
parameters under the reference's names (nn.Conv2d / nn.BatchNorm2d leaves, registered in the reference's order
so that ``isinstance(m, nn.Conv2d)`` scans such as BinOp / QuanOp pick the same target convs); the leaves'
own forward is never called.
"""
import math

import torch
import torch.nn as nn

from ..engine import Engine
from ..plan import Plan


class _Node(nn.Module):
    """Name-space node of the parameter tree."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder; call the CU-Net module itself")


class _HeadsFn(torch.autograd.Function):
    """Autograd bridge used when the caller drives the step with torch (loss.backward(); optimizer.step()),
    exactly like cu-net.py:171-183 does."""

    @staticmethod
    def forward(ctx, net, img, *params):
        eng = net._engine_for(img)
        eng.img.copy_(img)
        eng.forward(train=True)
        ctx.net, ctx.eng = net, eng
        return tuple(o.clone() for o in eng.head_outputs())

    @staticmethod
    def backward(ctx, *gouts):
        eng, net = ctx.eng, ctx.net
        p = eng.plan
        for h, g in zip(p.heads, gouts):
            buf = eng.G[h.name].view(eng.N, p.out_res, p.out_res, p.head_pad)
            buf.zero_()
            if g is not None:
                buf[..., :p.class_num].copy_(g.permute(0, 2, 3, 1))
        eng.backward()
        grads = []
        base = eng.grads.data_ptr()
        for name, prm in net._param_items:
            o, n, shape = eng.p_off[name]
            if prm.grad is not None and prm.grad.data_ptr() == base + 4 * o:
                prm.grad = None      # .grad was aliased to the flat bucket by bind_grads(): hand autograd a fresh tensor
            grads.append(eng.grads[o:o + n].view(shape).clone())
        return (None, None) + tuple(grads)


class CUNetB200(nn.Module):
    def __init__(self, init_chan_num, neck_size, growth_rate, class_num, layer_num, order, loss_num,
                 dtype="fp32", in_res=256, bits_i=0):
        super(CUNetB200, self).__init__()
        self.plan = Plan(class_num, layer_num, order, loss_num, neck_size, growth_rate, init_chan_num, in_res,
                         quan_input_bits=bits_i)
        self.loss_anchors = list(self.plan.anchors)
        self.layer_num = layer_num
        self.compute_dtype = dtype
        self._engines = {}
        self._store = None
        self._param_items = []
        self._build_tree()

    # -- parameter tree ------------------------------------------------------------------------
    def _leaf_parent(self, dotted):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        return node, parts[-1]

    def _build_tree(self):
        """Create nn.Conv2d / nn.BatchNorm2d leaves in the reference's registration order and initialise them
        with the reference's distributions (models/cu_net.py:322-334)."""
        convs, bns = {}, {}
        for s in self.plan.params:
            prefix = s.name.rsplit(".", 1)[0]
            if s.kind == "conv":
                co, ci, k, _ = s.shape
                parent, leaf = self._leaf_parent(prefix)
                m = nn.Conv2d(ci, co, k, stride=2 if k == 7 else 1, padding=k // 2, bias=False)
                stdv = 1.0 / math.sqrt(k * k * ci)
                m.weight.data.uniform_(-stdv, stdv)
                parent.add_module(leaf, m)
                convs[prefix] = m
            elif s.kind == "bn_weight":
                parent, leaf = self._leaf_parent(prefix)
                m = nn.BatchNorm2d(s.shape[0])
                m.weight.data.uniform_()
                m.bias.data.zero_()
                parent.add_module(leaf, m)
                bns[prefix] = m
        self._convs, self._bns = convs, bns

    def _bind(self, eng):
        """Move parameter / buffer storage into the engine's flat device buffers (views, no copies later)."""
        for s in self.plan.params:
            prefix, leaf = s.name.rsplit(".", 1)
            mod = self._convs.get(prefix) or self._bns.get(prefix)
            if s.kind in ("conv", "bn_weight", "bn_bias"):
                o, n, shape = eng.p_off[s.name]
                view = eng.params[o:o + n].view(shape)
                old = getattr(mod, leaf)
                view.copy_(old.data)
                prm = nn.Parameter(view, requires_grad=True)
                setattr(mod, leaf, prm)
            elif s.kind in ("bn_mean", "bn_var"):
                o, n, shape = eng.b_off[s.name]
                view = eng.bnbuf[o:o + n].view(shape)
                view.copy_(getattr(mod, leaf))
                setattr(mod, leaf, view)
            else:
                view = eng.counters[eng.cnt_idx[s.name]]
                view.copy_(getattr(mod, leaf))
                setattr(mod, leaf, view)
        self._param_items = [(n, p) for n, p in self.named_parameters()]

    def _engine_for(self, img):
        if not img.is_cuda:
            raise RuntimeError("cunet_b200 runs on CUDA tensors only (no CPU fallback); move the input with .cuda()")
        key = (int(img.shape[0]), img.device.index)
        eng = self._engines.get(key)
        if eng is None:
            # another batch size shares the parameter storage of the first engine
            eng = Engine(self.plan, img.shape[0], self.compute_dtype, img.device, share=self._store)
            if self._store is None:
                self._bind(eng)
                self._store = eng
            self._engines[key] = eng
        return eng

    def bind_grads(self):
        """Alias every parameter's ``.grad`` to its slice of the engine's flat gradient bucket (the fused Trainer
        path: BinOp / QuanOp and torch optimizers then see the gradients the backward kernels wrote)."""
        eng = self._store
        for name, prm in self._param_items:
            o, n, shape = eng.p_off[name]
            prm.grad = eng.grads[o:o + n].view(shape)

    def cuda(self, device=None):
        # parameters live in the engine's device buffers once the first engine exists
        if self._store is None:
            return super(CUNetB200, self).cuda(device)
        return self

    # -- reference API ------------------------------------------------------------------------
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.plan.in_res or x.shape[3] != self.plan.in_res:
            raise ValueError("expected input [N,3,%d,%d]" % (self.plan.in_res, self.plan.in_res))
        x = x.float()
        eng = self._engine_for(x)
        if self.training and torch.is_grad_enabled():
            return list(_HeadsFn.apply(self, x, *[p for _, p in self._param_items]))
        eng.img.copy_(x)
        eng.forward(train=self.training)
        return [o.clone() for o in eng.head_outputs()]

    def engine(self, batch, device=None):
        """The fused-step engine for a batch size (creates it on first use)."""
        dev = torch.device(device if device is not None else "cuda")
        dummy = torch.empty(batch, 0, device=dev)
        return self._engine_for(dummy)


def create_cu_net(neck_size, growth_rate, init_chan_num, class_num, layer_num, order, loss_num,
                  dtype="fp32", in_res=256, bits_i=0):
    """models/cu_net.py:362-368 (extra keyword arguments select the compute dtype / input size).

    bits_i > 0 builds the activation-quantized variant of models/cu_net_prev_version_wig.py: a QuanInput2d
    (utils/quantize.py:47-73, bitsI bits) in front of every dense-layer 3x3 conv and every head conv (:96-98,277-279),
    applied inside the fused convs' operand transform; together with QuanOp on the weights and gradients this is
    the training step of cu-net-prev-version-wig.py."""
    return CUNetB200(init_chan_num=init_chan_num, neck_size=neck_size, growth_rate=growth_rate,
                     class_num=class_num, layer_num=layer_num, order=order, loss_num=loss_num,
                     dtype=dtype, in_res=in_res, bits_i=bits_i)
