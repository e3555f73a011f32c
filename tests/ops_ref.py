"""Plain-PyTorch fp32 statements of each CUDA kernel's contract (test infrastructure).

Everything works on NHWC "pixel row" tensors exactly like the kernels do, so the same functions are
used (a) on the GPU to check single kernels and (b) on the CPU to check the op plan / backward
schedule against the oracle's autograd without a GPU.
"""
import torch
import torch.nn.functional as F


def rows_to_nchw(x, n, h, w):
    return x.float().reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


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
    return scale.float(), shift.float(), mean, var


def conv_fwd_ref(srcs, ups, n, h, w, scale, shift, weight, pool=False):
    """srcs: list of row tensors (at h,w or h/2,w/2 when up). Returns (out_rows, pool_idx or None)."""
    xs = []
    for s, up in zip(srcs, ups):
        if up:
            x = rows_to_nchw(s, n, h // 2, w // 2)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        else:
            x = rows_to_nchw(s, n, h, w)
        xs.append(x)
    x = torch.cat(xs, 1)
    a = F.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    y = F.conv2d(a, weight, padding=weight.shape[-1] // 2)
    idx = None
    if pool:
        y, flat = F.max_pool2d(y, 2, 2, return_indices=True)
        hh, ww = torch.div(flat, w, rounding_mode="floor"), flat % w
        idx = ((hh % 2) * 2 + (ww % 2)).to(torch.uint8)
        idx = nchw_to_rows(idx)
    return nchw_to_rows(y), idx
