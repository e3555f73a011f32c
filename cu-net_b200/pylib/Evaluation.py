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


MPII_PCKH_JOINTS = (0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15)     # pylib/Evaluation.py:92


def inverse_crop_transforms(center, scale, rot, res, size=200):
    """Batched inverse of pylib/HumanAug.py:10-34 (GetTransform): [N,3,3] float64 on the inputs' device."""
    center, scale, rot = center.double(), scale.double().reshape(-1), rot.double().reshape(-1)
    n = center.shape[0]
    h = size * scale
    t = torch.zeros(n, 3, 3, dtype=torch.float64, device=center.device)
    t[:, 0, 0] = res / h
    t[:, 1, 1] = res / h
    t[:, 0, 2] = res * (-center[:, 0] / h + 0.5)
    t[:, 1, 2] = res * (-center[:, 1] / h + 0.5)
    t[:, 2, 2] = 1
    rad = -rot * (torch.pi / 180)                      # "to match direction of rotation from cropping"
    sn, cs = torch.sin(rad), torch.cos(rad)
    rm = torch.zeros_like(t)
    rm[:, 0, 0], rm[:, 0, 1], rm[:, 1, 0], rm[:, 1, 1], rm[:, 2, 2] = cs, -sn, sn, cs, 1
    tm = torch.eye(3, dtype=torch.float64, device=center.device).repeat(n, 1, 1)
    tm[:, 0, 2] = -res / 2
    tm[:, 1, 2] = -res / 2
    ti = tm.clone()
    ti[:, :2, 2] *= -1
    return torch.linalg.inv(ti @ rm @ tm @ t)          # rot == 0: the rotation factors cancel to the identity


def final_preds_from_coords(output, coords, center, scale, res, rot):
    """The arithmetic of ``final_preds`` after the decode (device-agnostic, no per-joint Python loop):
    quarter-pixel refinement (pylib/Evaluation.py:113-121), +0.5, inverse crop transform of the 1-based points and
    truncation to int (:134-150, :179-187)."""
    n, c, hh, ww = output.shape
    coords = coords.float().clone()
    px = torch.floor(coords[..., 0]).long()
    py = torch.floor(coords[..., 1]).long()
    ok = (px > 1) & (px < res[0]) & (py > 1) & (py < res[1])
    flat = output.reshape(n, c, hh * ww).float()

    def at(y, x):
        idx = (y.clamp(0, hh - 1) * ww + x.clamp(0, ww - 1)).unsqueeze(-1)
        return flat.gather(2, idx).squeeze(-1)
    dx = at(py - 1, px) - at(py - 1, px - 2)
    dy = at(py, px - 1) - at(py - 2, px - 1)
    step = torch.stack([torch.sign(dx), torch.sign(dy)], -1) * 0.25
    coords = coords + torch.where(ok.unsqueeze(-1), step, torch.zeros_like(step)) + 0.5
    tinv = inverse_crop_transforms(center, scale, rot, res[0])
    # Evaluation.py's own TransformPts (:179-187) treats the points as 1-based: pts - 1 in, int() + 1 out
    homog = torch.cat([coords.double() - 1, torch.ones(n, c, 1, dtype=torch.float64, device=coords.device)], -1)
    out = torch.einsum("nij,ncj->nci", tinv, homog)[..., :2]
    return (torch.trunc(out) + 1).float()


def final_preds(output, center, scale, res, rot):
    """pylib/Evaluation.py:108-132: landmark predictions in original-image coordinates (decode on the GPU)."""
    dev = output.device
    return final_preds_from_coords(output, get_preds(output), center.to(dev), scale.to(dev), res, rot.to(dev))


def accuracy_origin_res(output, center, scale, res, grnd_pts, normalizers, rot):
    """pylib/Evaluation.py:88-106: PCKh at the original resolution over the 14 MPII joints."""
    dev = output.device
    pred = final_preds(output, center, scale, res, rot)
    dists = calc_dists(pred, grnd_pts.to(dev), normalizers.to(dev), use_zero=True)
    acc = torch.zeros(len(MPII_PCKH_JOINTS) + 1, device=dev)
    avg, cnt = 0.0, 0
    for i, cidx in enumerate(MPII_PCKH_JOINTS):
        acc[i + 1] = dist_acc(dists[cidx])
        if float(acc[i + 1]) >= 0:
            avg, cnt = avg + float(acc[i + 1]), cnt + 1
    if cnt:
        acc[0] = avg / cnt
    return acc
