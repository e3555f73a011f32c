"""Drop-in for the reference's ``utils/quantize.py`` (QuanOp, QuanInput2d) and for ``BinOp`` of
``models/cu_net_prev_version.py:17-92``, running on the multi-tensor CUDA kernels of csrc/quant.cu.

Call protocol (identical to cu-net-prev-version-wig.py:165,189-190 / cu-net-prev-version-bin.py:165,189-190,230,285):

    op = QuanOp(model)            # or BinOp(model)
    op.quantization()             # BinOp: op.binarization()
    ... forward / backward ...
    op.restore()
    op.updateQuanGradWeight()     # BinOp: op.updateBinaryGradWeight()
    optimizer.step()

Differences from the reference, on purpose: the bit widths are constructor arguments (defaults = the reference's
flag defaults bits_w=1, bits_i=8, bits_g=8, options/train_options.py:33-38) instead of an argparse run at import
time (utils/quantize.py:8-11).

Target selection is the reference's rule -- nn.Conv2d modules with modules()-index 1 .. count-2
(utils/quantize.py:80-102, models/cu_net_prev_version.py:19-41) -- applied to the module tree the reference applies it
to.  Both quantizers are only ever built on the *prev-version* models (cu-net-prev-version-bin.py:50,65,
cu-net-prev-version-wig.py:50,65), where the bottleneck 1x1 and every adapter are ``_EfficientDensenetBottleneck``
modules holding bare Parameters (cu_net_prev_version.py:118-157,166,217-231,294), so the nn.Conv2d set there is conv0
(:451), the dense-layer 3x3 ``conv.2`` (:170) and the ``layer_num`` heads (:348,463-466): for CU-Net-8 the rule selects
72 x [32,128,3,3] + 7 x [C,128,1,1].  The drop-in tree of models/cu_net.py registers every 1x1 as nn.Conv2d, so the
bare rule would select ~262 tensors there and train a different network; ``targets="prev_version"`` (the default for
a CU-Net drop-in module) reproduces the reference's set, ``targets="all_conv2d"`` is the bare rule (the default for any
other nn.Module).  tests/golden/binop_targets.pt pins the set against the REAL BinOp constructor.
"""
import ctypes as C

import torch
import torch.nn as nn

from .. import lib as L


def prev_version_conv_names(model):
    """Names of the drop-in tree's convs that are nn.Conv2d in the prev-version models, in modules() order."""
    names = []
    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d) and (name == "features.conv0" or name.endswith(".conv2")
                                         or (name.startswith("linears.") and name.endswith(".conv"))):
            names.append(name)
    return names


def target_names(model, targets="auto"):
    """Module names of the quantizer's target convs: modules()-index 1 .. count-2 of the selected Conv2d set."""
    if targets == "auto":
        targets = "prev_version" if hasattr(model, "plan") and hasattr(model, "loss_anchors") else "all_conv2d"
    if targets == "prev_version":
        names = prev_version_conv_names(model)
    elif targets == "all_conv2d":
        names = [n for n, m in model.named_modules() if isinstance(m, nn.Conv2d)]
    else:
        raise ValueError("targets must be 'auto', 'prev_version' or 'all_conv2d'")
    return names[1:len(names) - 1]          # leave out the first and the last Conv2d


def _targets(model, targets="auto"):
    mods = dict(model.named_modules())
    return [mods[n] for n in target_names(model, targets)]


class _MultiTensorOp(object):
    def __init__(self, model, targets="auto"):
        L.load()
        self.target_modules = [m.weight for m in _targets(model, targets)]
        self.num_of_params = len(self.target_modules)
        for w in self.target_modules:
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()):
                raise L.CunetError("quantizer targets must be contiguous fp32 CUDA parameters (no CPU fallback)")
        self.saved_params = [torch.empty_like(w.data) for w in self.target_modules]
        self._descs_dev = None
        self._grad_ptrs = None

    def _descs(self):
        grad_ptrs = tuple(w.grad.data_ptr() if w.grad is not None else 0 for w in self.target_modules)
        if self._descs_dev is None or grad_ptrs != self._grad_ptrs:
            structs, first = [], 0
            for w, s, gp in zip(self.target_modules, self.saved_params, grad_ptrs):
                co, ci, kh, kw = w.shape
                structs.append(L.QuantDesc(w.data.data_ptr(), s.data_ptr(), gp or None, co, ci, kh * kw, first))
                first += co
            raw = b"".join(bytes(x) for x in structs)
            self._descs_dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.target_modules[0].device)
            self._nblocks, self._grad_ptrs = first, grad_ptrs
        return C.c_void_p(self._descs_dev.data_ptr()), len(self.target_modules), self._nblocks

    def restore(self):
        if not self.num_of_params:
            return
        d, n, nb = self._descs()
        L.check(L.load().cunet_quant_restore(d, n, nb, L.stream_ptr()), "cunet_quant_restore")


class QuanOp(_MultiTensorOp):
    """utils/quantize.py:77-175."""

    def __init__(self, model, bits_w=1, bits_g=8, targets="auto"):
        super(QuanOp, self).__init__(model, targets)
        self.bits_w, self.bits_g = int(bits_w), int(bits_g)

    def quantization(self):
        if not self.num_of_params:
            return
        d, n, nb = self._descs()
        L.check(L.load().cunet_quant_forward(d, n, nb, 0, self.bits_w, self.bits_g, L.stream_ptr()),
                "cunet_quant_forward")

    def updateQuanGradWeight(self):
        if not self.num_of_params:
            return
        d, n, nb = self._descs()
        L.check(L.load().cunet_quant_grad(d, n, nb, 0, self.bits_w, self.bits_g, L.stream_ptr()), "cunet_quant_grad")


class BinOp(_MultiTensorOp):
    """models/cu_net_prev_version.py:17-92."""

    def binarization(self):
        if not self.num_of_params:
            return
        d, n, nb = self._descs()
        L.check(L.load().cunet_quant_forward(d, n, nb, 1, 1, 32, L.stream_ptr()), "cunet_quant_forward")

    def updateBinaryGradWeight(self):
        if not self.num_of_params:
            return
        d, n, nb = self._descs()
        L.check(L.load().cunet_quant_grad(d, n, nb, 1, 1, 32, L.stream_ptr()), "cunet_quant_grad")


class _QuanInputFn(torch.autograd.Function):
    """utils/quantize.py:47-63 (the reference's legacy non-static Function does not run on modern torch)."""

    @staticmethod
    def forward(ctx, x, bits):
        x = x.contiguous().float()
        y = torch.empty_like(x)
        L.check(L.load().cunet_quant_input_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_long(x.numel()),
                                               int(bits), L.stream_ptr()), "cunet_quant_input_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = gy.contiguous().float()
        gx = torch.empty_like(gy)
        L.check(L.load().cunet_quant_input_bwd(C.c_void_p(x.data_ptr()), C.c_void_p(gy.data_ptr()),
                                               C.c_void_p(gx.data_ptr()), C.c_long(x.numel()), L.stream_ptr()),
                "cunet_quant_input_bwd")
        return gx, None


class QuanInput2d(nn.Module):
    """utils/quantize.py:66-73."""

    def __init__(self, bits_i=8):
        super(QuanInput2d, self).__init__()
        self.layer_type = "QuanInput2d"
        self.bits_i = bits_i

    def forward(self, x):
        if not x.is_cuda:
            raise L.CunetError("QuanInput2d runs on CUDA tensors only (no CPU fallback)")
        return _QuanInputFn.apply(x, self.bits_i)
