"""CPU restatement of the weight/activation quantizers.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows
  * ``utils/quantize.py:15-42``   S / C / Q helpers
  * ``utils/quantize.py:47-63``   QuanInput (legacy autograd.Function -> restated as fwd/bwd pair)
  * ``utils/quantize.py:77-175``  QuanOp
  * ``models/cu_net_prev_version.py:17-92``  BinOp (torch-0.1.12 reductions keep dims, spelled
    out as ``keepdim=True`` exactly like utils/quantize.py:113,130-131,162-163,169-170)

All functions operate on lists of weight tensors [Cout, Cin, kh, kw] so they can be checked
both against the real ``utils/quantize.py`` (build container) and against the CUDA kernels.
"""
import torch


def S(bits):                                   # utils/quantize.py:15-16
    return 2.0 ** (bits - 1)


def C(x, bits=32):                             # utils/quantize.py:20-28
    if bits > 15 or bits == 1 or bits == 2:
        delta = 0.0
    else:
        delta = 1.0 / S(bits)
    return torch.clamp(x, -1 + delta, +1 - delta)


def Q(x, bits):                                # utils/quantize.py:33-42
    if bits > 15:
        return x
    if bits == 1:
        return torch.sign(x)
    if bits == 2:
        return torch.round(x)
    sc = S(bits)
    return torch.round(x * sc) / sc


def target_indices(num_conv):
    """Conv2d ``modules()`` indices 1 .. count-2 (utils/quantize.py:80-102,
    models/cu_net_prev_version.py:19-41); numpy.linspace(start,end,end-start+1) of the source."""
    return list(range(1, num_conv - 2 + 1))


def filter_mean_abs(w):
    """``w.norm(1,3,True).sum(2,True).sum(1,True).div(n)`` (utils/quantize.py:130-131)."""
    n = w[0].nelement()
    return w.abs().sum(dim=(1, 2, 3), keepdim=True) / n


# ------------------------------------------------------------------------------ QuanOp
def quanop_quantization(weights, bits_w, bits_g):
    """QuanOp.quantization (utils/quantize.py:104-149).  Returns (quantized, saved) lists."""
    out, saved = [], []
    for w in weights:
        w = w - w.mean(1, True)                                   # :110-115
        w = C(w, bits_g)                                          # :117-119 (uses bitsG!)
        saved.append(Q(w, bits_g))                                # :121-123
        if bits_w == 1:                                           # :127-134
            m = Q(filter_mean_abs(w).expand_as(w), bits_g)
            w = w.sign() * m
        if bits_w == 2:                                           # :135-147
            d = filter_mean_abs(w) * 0.7
            w = (w > d).float() - (w < -d).float()
        else:                                                     # :148-149 (also hit by bitsW==1)
            w = Q(C(w, bits_w), bits_w)
        out.append(w)
    return out, saved


def quanop_update_grad(weights, grads, bits_w, bits_g):
    """QuanOp.updateQuanGradWeight (utils/quantize.py:156-175); ``weights`` are the restored ones."""
    out = []
    for w, g in zip(weights, grads):
        if bits_w == 1:
            n = w[0].nelement()
            m = filter_mean_abs(w).expand_as(w).clone()
            m[w.lt(-1.0)] = 0
            m[w.gt(1.0)] = 0
            m = Q(m, bits_g)
            m = m * g
            m_add = (w.sign() * g).sum(dim=(1, 2, 3), keepdim=True) / n
            m_add = m_add.expand_as(w) * w.sign()
            g = (m + m_add) * (1.0 - 1.0 / w.size(1)) * n
        out.append(Q(C(g, bits_g), bits_g))
    return out


def quan_input_forward(x, bits_i):             # utils/quantize.py:52-55
    return Q(C(x, bits_i), bits_i)


def quan_input_backward(x, grad_out):          # utils/quantize.py:58-63
    g = grad_out.clone()
    g[x.ge(1)] = 0
    g[x.le(-1)] = 0
    return g


# ------------------------------------------------------------------------------ BinOp
def binop_binarization(weights):
    """BinOp.binarization (models/cu_net_prev_version.py:43-72).  Returns (binarized, saved)."""
    out, saved = [], []
    for w in weights:
        w = w - w.mean(1, True)                                   # :49-54
        w = w.clamp(-1.0, 1.0)                                    # :56-59
        saved.append(w.clone())                                   # :61-63 (not quantized)
        out.append(w.sign() * filter_mean_abs(w))                 # :65-72 (scale kept)
    return out, saved


def binop_update_grad(weights, grads):
    """BinOp.updateBinaryGradWeight (models/cu_net_prev_version.py:78-92)."""
    out = []
    for w, g in zip(weights, grads):
        n = w[0].nelement()
        m = filter_mean_abs(w).expand_as(w).clone()
        m[w.lt(-1.0)] = 0
        m[w.gt(1.0)] = 0
        m = m * g
        m_add = (w.sign() * g).sum(dim=(1, 2, 3), keepdim=True) / n
        m_add = m_add.expand_as(w) * w.sign()
        out.append((m + m_add) * (1.0 - 1.0 / w.size(1)) * n)
    return out
