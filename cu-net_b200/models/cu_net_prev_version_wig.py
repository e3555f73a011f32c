"""Name-compatible home of the activation-quantized ("wig": weights, inputs, gradients) CU-Net of the reference,
``models/cu_net_prev_version_wig.py::create_cu_net`` as imported by cu-net-prev-version-wig.py:23.

The reference file is the prev-version network (same graph as models/cu_net.py) with a ``QuanInput2d`` in front of
every dense-layer 3x3 conv (:96-98) and every head conv (:277-279) whenever ``bitsI <= 15``.  Here that is the same
fused-kernel network with ``bits_i`` set: the quantizer runs inside the convs' operand transform
(csrc/loaders.cuh::ActQuant), its straight-through backward inside the dgrad epilogues.  ``bits_i`` is a keyword
argument (default: the reference's flag default 8, options/train_options.py:35) instead of an import-time argparse."""
from .cu_net import create_cu_net as _create


def create_cu_net(neck_size, growth_rate, init_chan_num, class_num, layer_num, order, loss_num, dtype="fp32",
                  in_res=256, bits_i=8):
    return _create(neck_size, growth_rate, init_chan_num, class_num, layer_num, order, loss_num, dtype=dtype,
                   in_res=in_res, bits_i=bits_i if bits_i <= 15 else 0)
