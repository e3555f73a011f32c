"""Drop-in for ``pylib/Evaluation.get_preds`` (pylib/Evaluation.py:6-23): heatmap -> landmark decode on the GPU,
and for the heat-map-resolution PCK of the validation loop, ``Evaluation.accuracy`` (pylib/Evaluation.py:25-85).

``get_preds(scores[N,C,H,W]) -> float[N,C,2]``: argmax over H*W (first maximum), 1-based (x, y), zero where the
maximum is <= 0.  Runs the fused decode kernel of csrc/loss.cu (the same one the fused training step uses)."""
import ctypes as C

import torch

from .. import lib as L


def get_preds(scores):
    assert scores.dim() == 4, "Score maps should be 4-dim"
    if not scores.is_cuda:
        raise L.CunetError("cunet_b200 get_preds runs on CUDA tensors only (no CPU fallback)")
    n, c, h, w = scores.shape
    rows = scores.permute(0, 2, 3, 1).contiguous().float()            # NHWC, ld = C
    keys = torch.zeros(n * c, dtype=torch.int64, device=scores.device)
    loss = torch.zeros(2, dtype=torch.float64, device=scores.device)
    preds = torch.empty(n, c, 2, device=scores.device)
    mp = L.MseParams()
    mp.heads[0] = rows.data_ptr()
    mp.nheads = 1
    mp.target = scores.contiguous().float().data_ptr()                # unused for the decode (loss is discarded)
    mp.N, mp.C, mp.H, mp.W, mp.ld = n, c, h, w, c
    mp.loss, mp.keys, mp.grad_scale, mp.dtype = loss.data_ptr(), keys.data_ptr(), 1.0, L.F32
    lib = L.load()
    L.check(lib.cunet_mse_decode(C.byref(mp), L.stream_ptr()), "cunet_mse_decode")
    L.check(lib.cunet_decode_finalize(C.c_void_p(keys.data_ptr()), C.c_void_p(preds.data_ptr()), n * c, w,
                                      L.stream_ptr()), "cunet_decode_finalize")
    return preds


def calc_dists(preds, target, normalize, use_zero=False):
    """pylib/Evaluation.py:25-40, vectorised: [C, N] distances pred <-> target divided by normalize[n], -1 where the
    target coordinate is not above the boundary (1, or 0 with use_zero)."""
    preds, target = preds.float(), target.float()
    boundary = 0.0 if use_zero else 1.0
    valid = (target[..., 0] > boundary) & (target[..., 1] > boundary)                 # [N, C]
    d = torch.sqrt(((preds - target) ** 2).sum(-1)) / normalize.float().view(-1, 1)     # torch.dist = 2-norm
    return torch.where(valid, d, torch.full_like(d, -1.0)).t().contiguous()


def dist_acc(dists, thr=0.5):
    """pylib/Evaluation.py:42-53: fraction of the valid (!= -1) distances that are <= thr; -1 when none is valid."""
    valid = dists.ne(-1)
    nvalid = valid.sum()
    if int(nvalid) == 0:
        return torch.tensor(-1.0, device=dists.device)
    return (dists.le(thr) & valid).sum().float() / nvalid.float()


def accuracy_from_preds(preds, gts, width, idxs, thr=0.5):
    """The arithmetic of ``accuracy`` after the two decodes (device-agnostic)."""
    norm = torch.ones(preds.shape[0], device=preds.device) * width / 10
    dists = calc_dists(preds, gts, norm)
    acc = torch.zeros(len(idxs) + 1, device=preds.device)
    avg, cnt = 0.0, 0
    for i, c in enumerate(idxs):
        acc[i + 1] = dist_acc(dists[int(c)], thr)
        if float(acc[i + 1]) >= 0:
            avg, cnt = avg + float(acc[i + 1]), cnt + 1
    if cnt:
        acc[0] = avg / cnt
    return acc


def accuracy(output, target, idxs, thr=0.5):
    """PCK on the heat maps (pylib/Evaluation.py:55-85): both decodes run the fused CUDA kernel; acc[0] is the mean over
    ``idxs`` of the per-joint accuracies that have at least one valid target."""
    return accuracy_from_preds(get_preds(output), get_preds(target), output.shape[3], idxs, thr)
