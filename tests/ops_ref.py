"""Plain-PyTorch fp32 statements of each CUDA kernel's contract (test infrastructure).

Everything works on NHWC "pixel row" tensors exactly like the kernels do, so the same functions are
used (a) on the GPU to check single kernels and (b) on the CPU to check the op plan / backward
schedule against the oracle's autograd without a GPU.
"""
import torch
import torch.nn.functional as F


def rows_to_nchw(x, n, h, w, dtype=torch.float32):
    return x.to(dtype).reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


def nchw_to_rows(x):
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def tensor_stats(x):
    """[2*C] float64: per-channel sum and sum of squares of a row tensor."""
    xd = x.double()
    return torch.cat([xd.sum(0), (xd * xd).sum(0)])


def bn_coeffs(stats_list, counts, gamma, beta, eps=1e-5):
    """scale/shift of a train-mode BatchNorm over a virtual concat from per-source statistics."""
    means, vars_ = [], []
    for st, cnt in zip(stats_list, counts):
        c = st.numel() // 2
        m = st[:c] / cnt
        v = (st[c:] / cnt - m * m).clamp_min(0)
        means.append(m)
        vars_.append(v)
    mean, var = torch.cat(means), torch.cat(vars_)
    istd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.double() * istd
    shift = beta.double() - mean * scale
    return scale, shift, mean, var


def quan_act(a, bits):
    """QuanInput forward on an activated tensor (utils/quantize.py:52-55): Q(C(a, bits), bits)."""
    s = 2.0 ** (bits - 1)
    return torch.round(torch.clamp(a, -1 + 1 / s, 1 - 1 / s) * s) / s


def conv_fwd_ref(srcs, ups, n, h, w, scale, shift, weight, pool=False, dtype=torch.float32, act_bits=0,
                 identity=False):
    """srcs: list of row tensors (at h,w or h/2,w/2 when up). Returns (out_rows, pool_idx or None)."""
    xs = []
    for s, up in zip(srcs, ups):
        if up:
            x = rows_to_nchw(s, n, h // 2, w // 2, dtype)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        else:
            x = rows_to_nchw(s, n, h, w, dtype)
        xs.append(x)
    x = torch.cat(xs, 1)
    a = F.relu(x * scale.to(dtype).view(1, -1, 1, 1) + shift.to(dtype).view(1, -1, 1, 1))
    if identity:       # bn_train == 2: the stem's im2col operand, no BatchNorm and no ReLU in front of the conv
        a = x
    if act_bits:
        a = quan_act(a, act_bits)
    y = F.conv2d(a, weight.to(dtype), padding=weight.shape[-1] // 2)
    idx = None
    if pool:
        y, flat = F.max_pool2d(y, 2, 2, return_indices=True)
        hh, ww = torch.div(flat, w, rounding_mode="floor"), flat % w
        idx = ((hh % 2) * 2 + (ww % 2)).to(torch.uint8)
        idx = nchw_to_rows(idx)
    return nchw_to_rows(y), idx


def gstats_of(G, T, stats, count, eps=1e-5):
    """[2C] float64: sum G and sum G*xhat, xhat = (T - mean) * istd -- what the dgrad epilogue accumulates."""
    c = stats.numel() // 2
    mean = stats[:c] / count
    var = (stats[c:] / count - mean * mean).clamp_min(0)
    istd = 1.0 / torch.sqrt(var + eps)
    Gd, Td = G.double(), T.double()
    return torch.cat([Gd.sum(0), (Gd * (Td - mean) * istd).sum(0)])


def grad_coeffs(stats, gstats, count, eps=1e-5):
    """p, q, r of dT = p*G + q*T + r (batch-norm backward form, see include/cunet_b200.h)."""
    c = stats.numel() // 2
    mean = stats[:c] / count
    var = (stats[c:] / count - mean * mean).clamp_min(0)
    istd = 1.0 / torch.sqrt(var + eps)
    s1 = gstats[:c]
    s2 = gstats[c:]                      # sum G * xhat (accumulated centered)
    q = -istd * istd * s2 / count
    p = istd
    r = -istd * s1 / count - q * mean
    return p, q, r


def grad_src_eval(g, t, coeffs, pool_idx, n, h, w):
    """Full-resolution gradient rows dT [n*h*w][C] of a (possibly pooled) conv output.

    g, t: rows at h,w (or h/2,w/2 when pool_idx is given); coeffs: (p,q,r) or None (plain)."""
    dtype = torch.float64 if g.dtype == torch.float64 else torch.float32
    d = g.to(dtype)
    if coeffs is not None:
        p, q, r = coeffs
        d = p.to(dtype) * d + q.to(dtype) * t.to(dtype) + r.to(dtype)
    if pool_idx is None:
        return d
    c = d.shape[1]
    dn = rows_to_nchw(d, n, h // 2, w // 2, dtype)
    idx = pool_idx.reshape(n, h // 2, w // 2, c).permute(0, 3, 1, 2).long()
    full = torch.zeros(n, c, h, w, device=d.device, dtype=dtype)
    for pos in range(4):
        dy, dx = pos // 2, pos % 2
        full[:, :, dy::2, dx::2] = torch.where(idx == pos, dn, torch.zeros_like(dn))
    return nchw_to_rows(full)


def conv_bwd_ref(srcs, ups, n, h, w, scale, shift, mean, istd, gamma, weight, dy_rows, dtype=torch.float32,
                 act_bits=0):
    """Reference backward of cat->BN(train, stats held fixed)->ReLU->conv for the kernels' contract.

    Returns (G_contrib per source [at source resolution], dgamma, dbeta, dW).  G = gamma * dz."""
    xs = []
    for s, up in zip(srcs, ups):
        if up:
            x = F.interpolate(rows_to_nchw(s, n, h // 2, w // 2, dtype), scale_factor=2, mode="nearest")
        else:
            x = rows_to_nchw(s, n, h, w, dtype)
        xs.append(x)
    x = torch.cat(xs, 1)
    z = x * scale.to(dtype).view(1, -1, 1, 1) + shift.to(dtype).view(1, -1, 1, 1)
    a = F.relu(z)
    if act_bits:
        a = quan_act(a, act_bits)      # QuanInput: the conv (and its filter gradient) see the quantized activation
    a = a.detach().requires_grad_(True)
    wt = weight.detach().clone().to(dtype).requires_grad_(True)
    y = F.conv2d(a, wt, padding=weight.shape[-1] // 2)
    dy = rows_to_nchw(dy_rows, n, h, w, dtype)[:, :weight.shape[0]]
    da, dw = torch.autograd.grad(y, [a, wt], dy)
    dz = da * (z > 0).to(dtype)
    if act_bits:
        dz = dz * (z < 1).to(dtype)    # straight-through, zero where the activation is >= 1 (utils/quantize.py:58-63)
    xhat = (x - mean.view(1, -1, 1, 1).to(dtype)) * istd.view(1, -1, 1, 1).to(dtype)
    dbeta = dz.sum((0, 2, 3))
    dgamma = (dz * xhat).sum((0, 2, 3))
    gc = dz * gamma.to(dtype).view(1, -1, 1, 1)
    outs, c0 = [], 0
    for s, up in zip(srcs, ups):
        c = s.shape[1]
        part = gc[:, c0:c0 + c]
        if up:
            part = F.avg_pool2d(part, 2, 2) * 4
        outs.append(nchw_to_rows(part))
        c0 += c
    return outs, dgamma, dbeta, dw
