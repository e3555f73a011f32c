"""CPU restatement of the reference model ``models/cu_net.py`` (torch CPU fp32, autograd).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Written from the reference's behaviour, not
copied: a functional evaluator over a reference-named ``state_dict``.  Each function cites the
reference lines it follows (paths relative to /root/reference).

Parity pin: ``tests/test_oracle_golden.py`` checks this file against fixtures produced by the
real reference module (``oracle/gen_golden.py``) and, in the build container, against the real
module directly.
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default, models/cu_net.py:22,41,45,195,301
BN_MOMENTUM = 0.1


def loss_anchors(layer_num, loss_num):
    """models/cu_net.py:274-283."""
    assert 1 <= loss_num <= layer_num
    every = float(layer_num) / float(loss_num)
    anchors = []
    for i in range(loss_num):
        a = int(round(every * (i + 1)))
        if a <= layer_num:
            anchors.append(a)
    assert layer_num in anchors and len(anchors) == loss_num
    return anchors


def state_spec(class_num, layer_num, order, neck_size=4, growth_rate=32, init_chan_num=128):
    """(name, shape, kind) for every state_dict entry, in the reference's registration order.

    Follows the constructors: features models/cu_net.py:299-304; _CU_Net :228-250 (ModuleLists
    ``down_blocks``, ``up_blocks`` are attributes created before ``neck_block``); _DenseBlock
    :68-113 (layers, adapters_ahead, adapters_skip); heads :312-314; _IntermediaBlock :147-164.
    """
    L, K, g, C0 = layer_num, order, growth_rate, init_chan_num
    bott = neck_size * g
    spec = []

    def bn(prefix, c):
        spec.append((prefix + ".weight", (c,), "bn_weight"))
        spec.append((prefix + ".bias", (c,), "bn_bias"))
        spec.append((prefix + ".running_mean", (c,), "bn_mean"))
        spec.append((prefix + ".running_var", (c,), "bn_var"))
        spec.append((prefix + ".num_batches_tracked", (), "bn_count"))

    def conv(prefix, co, ci, k):
        spec.append((prefix + ".weight", (co, ci, k, k), "conv"))

    conv("features.conv0", C0, 3, 7)
    bn("features.norm0", C0)

    def dense_block(prefix, in_num, requires_skip, is_up):
        max_in = in_num + K * g
        for i in range(L):
            cin = in_num + i * g if i < K else max_in
            bn("%s.layers.%d.norm1" % (prefix, i), cin)
            conv("%s.layers.%d.conv1" % (prefix, i), bott, cin, 1)
            bn("%s.layers.%d.norm2" % (prefix, i), bott)
            conv("%s.layers.%d.conv2" % (prefix, i), g, bott, 3)
        out_num = in_num // 2 if is_up else in_num
        ad_in = [in_num + (i + 1) * g if i < K else max_in + g for i in range(L)]
        for i in range(L):
            bn("%s.adapters_ahead.%d.adapter_norm" % (prefix, i), ad_in[i])
            conv("%s.adapters_ahead.%d.adapter_conv" % (prefix, i), out_num, ad_in[i], 1)
        if requires_skip:
            for i in range(L):
                bn("%s.adapters_skip.%d.adapter_norm" % (prefix, i), ad_in[i])
                conv("%s.adapters_skip.%d.adapter_conv" % (prefix, i), out_num, ad_in[i], 1)

    for j in range(4):
        dense_block("hg.down_blocks.%d" % j, C0, True, False)
    for j in range(4):
        dense_block("hg.up_blocks.%d" % j, 2 * C0, False, True)
    dense_block("hg.neck_block", C0, False, False)
    for i in range(L):
        bn("linears.%d.norm" % i, C0)
        conv("linears.%d.conv" % i, class_num, C0, 1)
    max_in = C0 + K * C0
    for i in range(L - 1):
        cin = C0 + (i + 1) * C0 if i < K else max_in
        bn("intermedia.adapters.%d.adapter_norm" % i, cin)
        conv("intermedia.adapters.%d.adapter_conv" % i, C0, cin, 1)
    return spec


def init_state(class_num, layer_num, order, seed=0, **kw):
    """Seeded state_dict with the reference's init distributions (models/cu_net.py:322-334):
    conv ~ U(-1/sqrt(k*k*Cin), +), BN weight ~ U(0,1), BN bias 0; running stats (0, 1).
    (Own RNG stream -- not bit-identical to constructing the reference module.)"""
    gen = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape, kind in state_spec(class_num, layer_num, order, **kw):
        if kind == "conv":
            stdv = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            sd[name] = (torch.rand(shape, generator=gen) * 2 - 1) * stdv
        elif kind == "bn_weight":
            sd[name] = torch.rand(shape, generator=gen)
        elif kind == "bn_bias" or kind == "bn_mean":
            sd[name] = torch.zeros(shape)
        elif kind == "bn_var":
            sd[name] = torch.ones(shape)
        else:
            sd[name] = torch.zeros((), dtype=torch.long)
    return sd


class _QuanInputFn(torch.autograd.Function):
    """QuanInput of utils/quantize.py:47-63 as a modern static Function: forward Q(C(x, bitsI), bitsI) (:52-55),
    backward the incoming gradient with the entries where x >= 1 or x <= -1 zeroed (:58-63).  The arithmetic is pinned
    against the REAL forward / backward bodies by tests/golden/quaninput.pt (oracle/quantize_oracle.py)."""

    @staticmethod
    def forward(ctx, x, bits):
        from . import quantize_oracle
        ctx.save_for_backward(x)
        return quantize_oracle.quan_input_forward(x, bits)

    @staticmethod
    def backward(ctx, g):
        from . import quantize_oracle
        (x,) = ctx.saved_tensors
        return quantize_oracle.quan_input_backward(x, g), None


class OracleCUNet(object):
    """Functional evaluator of the reference network over a reference-named state_dict.

    quan_input_bits > 0: the activation-quantized variant (models/cu_net_prev_version_wig.py: QuanInput2d between
    relu.2 and conv.2 of every dense layer, :96-98, and between relu and conv of every head, :277-279)."""

    def __init__(self, state, class_num, layer_num, order, loss_num,
                 neck_size=4, growth_rate=32, init_chan_num=128, double_bn_update=True, quan_input_bits=0):
        self.quan_input_bits = quan_input_bits
        if order >= layer_num:                      # models/cu_net.py:285-287 (exit())
            raise SystemExit("order is larger than the layer number.")
        self.L, self.K, self.g = layer_num, order, growth_rate
        self.class_num = class_num
        self.anchors = loss_anchors(layer_num, loss_num)
        self.training = True
        # models/cu_net.py:30-33,58-61: BNs inside cp.checkpoint run forward twice per training
        # step when inputs require grad -> momentum applied twice.
        self.double_bn_update = double_bn_update
        self.state = OrderedDict()
        for k, v in state.items():
            if k.startswith("module."):
                k = k[7:]
            self.state[k] = v.detach().clone()
        self.param_names = [n for n, _, kind in state_spec(class_num, layer_num, order,
                                                           neck_size, growth_rate, init_chan_num)
                            if kind in ("conv", "bn_weight", "bn_bias")]
        for n in self.param_names:
            self.state[n] = self.state[n].float().requires_grad_(True)

    def parameters(self):
        return [self.state[n] for n in self.param_names]

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def zero_grad(self):
        for p in self.parameters():
            p.grad = None

    # -- primitives ------------------------------------------------------------------
    def _bn_relu(self, x, prefix, checkpointed):
        s = self.state
        w, b = s[prefix + ".weight"], s[prefix + ".bias"]
        rm, rv = s[prefix + ".running_mean"], s[prefix + ".running_var"]
        reps = 2 if (checkpointed and self.double_bn_update and self.training
                     and torch.is_grad_enabled()) else 1
        if self.training:
            for _ in range(reps - 1):   # extra statistics update from the checkpoint re-run
                with torch.no_grad():
                    F.batch_norm(x.detach(), rm, rv, None, None, True, BN_MOMENTUM, BN_EPS)
            key = prefix + ".num_batches_tracked"
            if key in s:
                s[key] = s[key] + reps
        y = F.batch_norm(x, rm, rv, w, b, self.training, BN_MOMENTUM, BN_EPS)
        return F.relu(y)

    def _cat_bn_relu_conv1x1(self, inputs, norm, conv, checkpointed=True):
        """_bn_function_factory, models/cu_net.py:11-17."""
        x = torch.cat(inputs, 1)
        return F.conv2d(self._bn_relu(x, norm, checkpointed), self.state[conv + ".weight"])

    def _dense_layer(self, inputs, prefix):
        """_DenseLayer.forward, models/cu_net.py:52-65 (drop_rate is always 0)."""
        bott = self._cat_bn_relu_conv1x1(inputs, prefix + ".norm1", prefix + ".conv1")
        y = self._bn_relu(bott, prefix + ".norm2", False)
        if self.quan_input_bits:                       # cu_net_prev_version_wig.py:96-98
            y = _QuanInputFn.apply(y, self.quan_input_bits)
        return F.conv2d(y, self.state[prefix + ".conv2.weight"], padding=1)

    def _dense_block(self, x, i, prefix, saved, requires_skip):
        """_DenseBlock.forward, models/cu_net.py:115-144.  ``saved`` is the block's FIFO."""
        if i == 0:
            del saved[:]
        xs = list(x) if isinstance(x, list) else [x]
        xs = xs + saved                                            # :127
        out = self._dense_layer(xs, "%s.layers.%d" % (prefix, i))  # :132
        if i < self.K:                                             # :133-137
            saved.append(out)
        elif len(saved) != 0:
            saved.pop(0)
            saved.append(out)
        xs.append(out)                                             # :138
        ahead = self._cat_bn_relu_conv1x1(
            xs, "%s.adapters_ahead.%d.adapter_norm" % (prefix, i),
            "%s.adapters_ahead.%d.adapter_conv" % (prefix, i))
        if requires_skip:
            skip = self._cat_bn_relu_conv1x1(
                xs, "%s.adapters_skip.%d.adapter_norm" % (prefix, i),
                "%s.adapters_skip.%d.adapter_conv" % (prefix, i))
            return ahead, skip
        return ahead

    def _hourglass(self, x, i, fifo):
        """_CU_Net.forward, models/cu_net.py:252-269."""
        skips = [None] * 4
        for j in range(4):
            x, skips[j] = self._dense_block(x, i, "hg.down_blocks.%d" % j, fifo["d%d" % j], True)
            x = F.max_pool2d(x, 2, 2)
        x = self._dense_block(x, i, "hg.neck_block", fifo["n"], False)
        for j in (3, 2, 1, 0):
            x = F.interpolate(x, scale_factor=2, mode="nearest")   # nn.Upsample default mode
            x = self._dense_block([x, skips[j]], i, "hg.up_blocks.%d" % j, fifo["u%d" % j], False)
        return x

    def _intermedia(self, x, i, saved):
        """_IntermediaBlock.forward, models/cu_net.py:166-190."""
        if i == 0:
            del saved[:]
            if self.K != 0:
                saved.append(x)
            return x
        xs = [x] + saved
        out = self._cat_bn_relu_conv1x1(
            xs, "intermedia.adapters.%d.adapter_norm" % (i - 1),
            "intermedia.adapters.%d.adapter_conv" % (i - 1))
        if i < self.K:
            saved.append(out)
        elif len(saved) != 0:
            saved.pop(0)
            saved.append(out)
        return out

    def forward(self, img):
        """_CU_Net_Wrapper.forward, models/cu_net.py:336-360."""
        s = self.state
        x = F.conv2d(img, s["features.conv0.weight"], stride=2, padding=3)
        x = self._bn_relu(x, "features.norm0", False)
        x = F.max_pool2d(x, 2, 2)
        fifo = {k: [] for k in ["d0", "d1", "d2", "d3", "u0", "u1", "u2", "u3", "n"]}
        inter = []
        outs = []
        for i in range(self.L):
            x = self._intermedia(x, i, inter)
            x = self._hourglass(x, i, fifo)
            if (i + 1) in self.anchors:
                y = self._bn_relu(x, "linears.%d.norm" % i, False)
                if self.quan_input_bits:               # cu_net_prev_version_wig.py:277-279
                    y = _QuanInputFn.apply(y, self.quan_input_bits)
                outs.append(F.conv2d(y, s["linears.%d.conv.weight" % i]))
        assert len(outs) == len(self.anchors)
        return outs

    __call__ = forward


def multi_loss_mse(outputs, heatmap):
    """cu-net.py:175-178."""
    loss = 0
    for o in outputs:
        d = (o - heatmap) ** 2
        loss = loss + d.sum() / d.numel()
    return loss


def rmsprop_step(params, grads, square_avg, lr, alpha=0.99, eps=1e-8):
    """torch.optim.RMSprop(lr, alpha=0.99, eps=1e-8, momentum=0, weight_decay=0), cu-net.py:60-61.
    v = alpha v + (1-alpha) g^2 ; p -= lr * g / (sqrt(v) + eps)."""
    with torch.no_grad():
        for p, g, v in zip(params, grads, square_avg):
            v.mul_(alpha).addcmul_(g, g, value=1 - alpha)
            p.addcdiv_(g, v.sqrt().add_(eps), value=-lr)
