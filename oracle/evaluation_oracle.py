"""CPU restatement of the heatmap decode (``pylib/Evaluation.py:6-23``).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference file cannot be imported
(``import HumanAug`` is a python-2 implicit relative import), so ``get_preds`` is restated.
"""
import torch


def get_preds(scores):
    """argmax over H*W (first maximum), 1-based (x, y), zeroed where max <= 0.

    x = idx % W + 1 (Evaluation.py:18), y = floor(idx / H) + 1 (Evaluation.py:19 divides by
    size(2): square maps only)."""
    assert scores.dim() == 4
    n, c, h, w = scores.shape
    maxval, idx = torch.max(scores.reshape(n, c, -1), 2)
    idx = idx.view(n, c, 1) + 1
    preds = idx.repeat(1, 1, 2).float()
    preds[:, :, 0] = (preds[:, :, 0] - 1) % w + 1
    preds[:, :, 1] = torch.floor((preds[:, :, 1] - 1) / h) + 1
    mask = maxval.view(n, c, 1).gt(0).repeat(1, 1, 2).float()
    return preds * mask


def calc_dists(preds, target, normalize, use_zero=False):
    """pylib/Evaluation.py:25-40, loop for loop."""
    preds, target, normalize = preds.float(), target.float(), normalize.float()
    dists = torch.zeros(preds.size(1), preds.size(0))
    boundary = 0 if use_zero else 1
    for n in range(preds.size(0)):
        for c in range(preds.size(1)):
            if target[n, c, 0] > boundary and target[n, c, 1] > boundary:
                dists[c, n] = torch.dist(preds[n, c, :], target[n, c, :]) / normalize[n]
            else:
                dists[c, n] = -1
    return dists


def dist_acc(dists, thr=0.5):
    """pylib/Evaluation.py:42-53."""
    if dists.ne(-1).sum() > 0:
        return dists.le(thr).eq(dists.ne(-1)).sum().float() / dists.ne(-1).sum().float()
    return -1


def accuracy(output, target, idxs, thr=0.5):
    """pylib/Evaluation.py:55-85."""
    preds, gts = get_preds(output), get_preds(target)
    norm = torch.ones(preds.size(0)) * output.size(3) / 10
    dists = calc_dists(preds, gts, norm)
    acc = torch.zeros(len(idxs) + 1)
    avg_acc, cnt = 0, 0
    for i in range(len(idxs)):
        acc[i + 1] = dist_acc(dists[idxs[i]])
        if acc[i + 1] >= 0:
            avg_acc = avg_acc + acc[i + 1]
            cnt += 1
    if cnt != 0:
        acc[0] = avg_acc / cnt
    return acc


def flip_channels(maps):
    """pylib/HumanAug.py:198-210 (numpy reversed view of the last axis)."""
    return torch.from_numpy(maps.numpy()[..., ::-1].copy()).float()


def shuffle_channels_for_horizontal_flipping(maps, flip_indxs):
    """pylib/HumanAug.py:177-196: in-place pairwise swaps along the channel axis."""
    dim = 1 if maps.ndimension() == 4 else 0
    for i in range(0, len(flip_indxs)):
        idx1, idx2 = flip_indxs[i]
        tmp = maps.narrow(dim, idx1, 1).clone()
        maps.narrow(dim, idx1, 1).copy_(maps.narrow(dim, idx2, 1))
        maps.narrow(dim, idx2, 1).copy_(tmp)
    return maps
