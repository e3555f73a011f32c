"""Static op plan of a CU-Net: the network compiled once into a list of fused-conv ops.

What the reference does dynamically -- one hourglass object re-entered ``layer_num`` times with per-block
FIFOs of the last ``order`` outputs (models/cu_net.py:115-144, 166-190, 252-269, 336-360) -- is unrolled here
into a static list of ops over named tensors.  A ``torch.cat`` never exists: every op carries the segment
table of its virtual concat, in the channel order the reference's weights expect:

    dense layer conv1 :  [X | U,S] + [O_{i-m} .. O_{i-1}]            (models/cu_net.py:127,132)
    adapters          :  the same + [O_i]                             (models/cu_net.py:138-142)
    intermedia        :  [x_i] + [z_{i-m} .. z_{i-1}]                 (models/cu_net.py:182-188)

Pure Python, no torch: the plan is data.  ``engine.py`` binds it to device buffers and kernels;
``tests/plan_emulator.py`` executes the same plan with plain PyTorch on the CPU to check it against the oracle.
"""
from collections import OrderedDict


def loss_anchors(layer_num, loss_num):
    """1-based U-Net indices that get a heatmap head (models/cu_net.py:274-283)."""
    if not (1 <= loss_num <= layer_num):
        raise AssertionError("need 1 <= loss_num <= layer_num")
    every = float(layer_num) / float(loss_num)
    anchors = [int(round(every * (k + 1))) for k in range(loss_num)]
    anchors = [a for a in anchors if a <= layer_num]
    assert layer_num in anchors and len(anchors) == loss_num
    return anchors


class ParamSpec(object):
    __slots__ = ("name", "shape", "kind", "offset", "numel")

    def __init__(self, name, shape, kind):
        self.name, self.shape, self.kind = name, tuple(shape), kind
        n = 1
        for s in shape:
            n *= s
        self.numel = n
        self.offset = -1


def param_specs(class_num, layer_num, order, neck_size=4, growth_rate=32, init_chan_num=128):
    """state_dict entries in the reference's registration order.

    kinds: conv | bn_weight | bn_bias | bn_mean | bn_var | bn_count."""
    L, K, g, C0 = layer_num, order, growth_rate, init_chan_num
    bott = neck_size * g
    out = []

    def bn(prefix, c):
        for suffix, kind in (("weight", "bn_weight"), ("bias", "bn_bias"), ("running_mean", "bn_mean"),
                             ("running_var", "bn_var")):
            out.append(ParamSpec("%s.%s" % (prefix, suffix), (c,), kind))
        out.append(ParamSpec(prefix + ".num_batches_tracked", (), "bn_count"))

    def conv(prefix, co, ci, k):
        out.append(ParamSpec(prefix + ".weight", (co, ci, k, k), "conv"))

    conv("features.conv0", C0, 3, 7)
    bn("features.norm0", C0)

    def block(prefix, in_num, skip, is_up):
        max_in = in_num + K * g
        for i in range(L):
            cin = in_num + min(i, K) * g
            bn("%s.layers.%d.norm1" % (prefix, i), cin)
            conv("%s.layers.%d.conv1" % (prefix, i), bott, cin, 1)
            bn("%s.layers.%d.norm2" % (prefix, i), bott)
            conv("%s.layers.%d.conv2" % (prefix, i), g, bott, 3)
        out_num = in_num // 2 if is_up else in_num
        for kind in (["adapters_ahead", "adapters_skip"] if skip else ["adapters_ahead"]):
            for i in range(L):
                cin = in_num + (min(i, K) + 1) * g
                bn("%s.%s.%d.adapter_norm" % (prefix, kind, i), cin)
                conv("%s.%s.%d.adapter_conv" % (prefix, kind, i), out_num, cin, 1)
        del max_in

    for j in range(4):
        block("hg.down_blocks.%d" % j, C0, True, False)
    for j in range(4):
        block("hg.up_blocks.%d" % j, 2 * C0, False, True)
    block("hg.neck_block", C0, False, False)
    for i in range(L):
        bn("linears.%d.norm" % i, C0)
        conv("linears.%d.conv" % i, class_num, C0, 1)
    for i in range(L - 1):
        cin = C0 * (1 + min(i + 1, K))
        bn("intermedia.adapters.%d.adapter_norm" % i, cin)
        conv("intermedia.adapters.%d.adapter_conv" % i, C0, cin, 1)
    return out


class TensorSpec(object):
    """An activation tensor in HBM: pixel rows [N*res*res][C]."""
    __slots__ = ("name", "C", "res", "fp32", "producer", "consumers", "index")

    def __init__(self, name, C, res, fp32=False):
        self.name, self.C, self.res, self.fp32 = name, C, res, fp32
        self.producer = None
        self.consumers = []      # (op, segment index) in forward order
        self.index = -1

    def __repr__(self):
        return "T(%s,%d@%d)" % (self.name, self.C, self.res)


class ConvOp(object):
    """One fused  cat -> BN -> ReLU -> conv [-> maxpool]  op."""
    __slots__ = ("name", "kind", "norm", "conv", "srcs", "out", "res", "taps", "cout", "cout_pad", "pool",
                 "checkpointed", "unet", "index")

    def __init__(self, name, kind, norm, conv, srcs, out, res, taps, cout, cout_pad, pool, checkpointed, unet):
        self.name, self.kind, self.norm, self.conv = name, kind, norm, conv
        self.srcs, self.out, self.res, self.taps = srcs, out, res, taps     # srcs: [(TensorSpec, up)]
        self.cout, self.cout_pad, self.pool = cout, cout_pad, pool
        self.checkpointed, self.unet = checkpointed, unet
        self.index = -1

    @property
    def cin(self):
        return sum(t.C for t, _ in self.srcs)

    @property
    def quan_input(self):
        """True for the convs that sit behind a QuanInput2d in the wig model: dense-layer conv2 and the heads."""
        return self.kind in ("3x3", "head")


class Plan(object):
    def __init__(self, class_num, layer_num, order, loss_num, neck_size=4, growth_rate=32, init_chan_num=128,
                 in_res=256, quan_input_bits=0):
        """quan_input_bits: 0, or the bit width of the QuanInput2d layers of the activation-quantized ("wig") model
        variant -- between relu.2 and conv.2 of every dense layer and between relu and conv of every head
        (models/cu_net_prev_version_wig.py:96-98, 277-279); the fused convs apply it inside their operand transform."""
        if quan_input_bits and not (2 <= quan_input_bits <= 8):
            raise ValueError("quan_input_bits must be 0 or in 2..8 (the bf16 operand carries 8 significant bits)")
        self.quan_input_bits = int(quan_input_bits)
        if order >= layer_num:
            # the reference prints and calls exit() (models/cu_net.py:285-287)
            raise SystemExit("order is larger than the layer number.")
        if neck_size * growth_rate != 128 or init_chan_num != 128 or growth_rate != 32:
            raise ValueError("the sm_100a kernels are built for the reference configuration "
                             "neck_size=4, growth_rate=32, init_chan_num=128 (cu-net.py:46)")
        if in_res % 64:
            raise ValueError("input resolution must be a multiple of 64")
        if order + 3 > 8:
            # widest virtual concat = an up-block adapter: [U, S, O_{i-K} .. O_i] = K + 3 segments (CUNET_MAX_SEG = 8)
            raise ValueError("order %d needs %d concat segments per conv; the kernels take at most 8 (order <= 5)"
                             % (order, order + 3))
        self.class_num, self.L, self.K, self.loss_num = class_num, layer_num, order, loss_num
        self.g, self.C0, self.bott = growth_rate, init_chan_num, neck_size * growth_rate
        self.in_res, self.stem_res, self.out_res = in_res, in_res // 2, in_res // 4
        self.anchors = loss_anchors(layer_num, loss_num)
        self.head_pad = ((class_num + 15) // 16) * 16
        self.params = param_specs(class_num, layer_num, order, neck_size, growth_rate, init_chan_num)
        self.tensors = OrderedDict()
        self.ops = []
        self.heads = []            # head output tensors, in order
        self._build()
        for i, t in enumerate(self.tensors.values()):
            t.index = i
        for i, op in enumerate(self.ops):
            op.index = i

    # ------------------------------------------------------------------------------------------
    def _tensor(self, name, C, res, fp32=False):
        t = TensorSpec(name, C, res, fp32)
        assert name not in self.tensors
        self.tensors[name] = t
        return t

    def _op(self, name, kind, norm, conv, srcs, out, res, taps, cout, cout_pad=None, pool=False,
            checkpointed=True, unet=0):
        op = ConvOp(name, kind, norm, conv, list(srcs), out, res, taps, cout, cout_pad or cout, pool,
                    checkpointed, unet)
        out.producer = op
        for si, (t, _) in enumerate(op.srcs):
            t.consumers.append((op, si))
        self.ops.append(op)
        return op

    def _dense_block(self, prefix, tag, i, xs, fifo, res, skip, pool_ahead):
        """_DenseBlock.forward (models/cu_net.py:115-144). xs: [(tensor, up)]. Returns (ahead, skip)."""
        if i == 0:
            del fifo[:]
        srcs = list(xs) + [(t, False) for t in fifo]
        lay = "%s.layers.%d" % (prefix, i)
        bott = self._tensor("%s.u%d.bott" % (tag, i), self.bott, res)
        self._op(lay + ".conv1", "1x1", lay + ".norm1", lay + ".conv1", srcs, bott, res, 1, self.bott, unet=i)
        o = self._tensor("%s.u%d.O" % (tag, i), self.g, res)
        self._op(lay + ".conv2", "3x3", lay + ".norm2", lay + ".conv2", [(bott, False)], o, res, 9, self.g,
                 checkpointed=False, unet=i)
        if i < self.K:
            fifo.append(o)
        elif len(fifo) != 0:
            fifo.pop(0)
            fifo.append(o)
        srcs2 = srcs + [(o, False)]
        ad = "%s.adapters_ahead.%d" % (prefix, i)
        ahead = self._tensor("%s.u%d.ahead" % (tag, i), self.C0, res // 2 if pool_ahead else res)
        self._op(ad, "1x1", ad + ".adapter_norm", ad + ".adapter_conv", srcs2, ahead, res, 1, self.C0,
                 pool=pool_ahead, unet=i)
        sk = None
        if skip:
            ad = "%s.adapters_skip.%d" % (prefix, i)
            sk = self._tensor("%s.u%d.skip" % (tag, i), self.C0, res)
            self._op(ad, "1x1", ad + ".adapter_norm", ad + ".adapter_conv", srcs2, sk, res, 1, self.C0, unet=i)
        return ahead, sk

    def _build(self):
        R = self.out_res
        self.stem_y = self._tensor("stem.y", self.C0, self.stem_res)        # conv0 output
        x = self._tensor("stem.x", self.C0, R)                              # after norm0/relu0/pool0
        self.stem_x = x
        fifo = {k: [] for k in ["d0", "d1", "d2", "d3", "u0", "u1", "u2", "u3", "n"]}
        inter = []
        for i in range(self.L):
            # _IntermediaBlock.forward (models/cu_net.py:166-190)
            if i == 0:
                if self.K != 0:
                    inter.append(x)
            else:
                srcs = [(x, False)] + [(t, False) for t in inter]
                z = self._tensor("inter.u%d" % i, self.C0, R)
                nm = "intermedia.adapters.%d" % (i - 1)
                self._op(nm, "1x1", nm + ".adapter_norm", nm + ".adapter_conv", srcs, z, R, 1, self.C0, unet=i)
                if i < self.K:
                    inter.append(z)
                elif len(inter) != 0:
                    inter.pop(0)
                    inter.append(z)
                x = z
            # _CU_Net.forward (models/cu_net.py:252-269)
            skips = [None] * 4
            for j in range(4):
                res = R >> j
                x, skips[j] = self._dense_block("hg.down_blocks.%d" % j, "d%d" % j, i, [(x, False)],
                                                fifo["d%d" % j], res, True, True)      # + maxpool (:260)
            x, _ = self._dense_block("hg.neck_block", "n", i, [(x, False)], fifo["n"], R >> 4, False, False)
            for j in (3, 2, 1, 0):
                res = R >> j
                x, _ = self._dense_block("hg.up_blocks.%d" % j, "u%d" % j, i, [(x, True), (skips[j], False)],
                                         fifo["u%d" % j], res, False, False)           # upsample (:265)
            if (i + 1) in self.anchors:
                nm = "linears.%d" % i
                h = self._tensor("head.u%d" % i, self.head_pad, R, fp32=True)
                self._op(nm, "head", nm + ".norm", nm + ".conv", [(x, False)], h, R, 1, self.class_num,
                         cout_pad=self.head_pad, checkpointed=False, unet=i)
                self.heads.append(h)
        assert len(self.heads) == len(self.anchors)

    # ------------------------------------------------------------------------------------------
    def backward_schedule(self):
        """Reverse op order with, per (op, segment): accumulate flag (False for the first writer of a
        tensor's gradient accumulator in backward order) and last flag (True for the last writer, which also
        produces the sums the producer's backward needs)."""
        seen = {}
        sched = []
        for op in reversed(self.ops):
            flags = []
            for t, _ in op.srcs:
                first = t.name not in seen
                seen[t.name] = op
                flags.append([not first, False])
            sched.append((op, flags))
        # last writer = the earliest consumer in forward order
        for op, flags in sched:
            for si, (t, _) in enumerate(op.srcs):
                if t.consumers[0][0] is op and t.consumers[0][1] == si:
                    flags[si][1] = True
        return sched

    def conv_flops_per_image(self):
        """2*MACs of all convolutions, forward."""
        f = 2.0 * 147 * self.C0 * self.stem_res * self.stem_res
        for op in self.ops:
            f += 2.0 * op.cin * op.taps * op.cout * op.res * op.res
        return f
