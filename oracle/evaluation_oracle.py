"""CPU restatement of the heatmap decode (``pylib/Evaluation.py:6-23``).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  The reference file cannot be imported
(``import HumanAug`` is a python-2 implicit relative import), so ``get_preds`` is restated.
"""
import math

import numpy as np
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


def get_transform(center, scale, rot, res, size):
    """pylib/Evaluation.py:152-177 (GetTransform; identical to pylib/HumanAug.py:10-34)."""
    h = size * scale
    t = np.zeros((3, 3))
    t[0, 0] = float(res) / h
    t[1, 1] = float(res) / h
    t[0, 2] = res * (-float(center[0]) / h + .5)
    t[1, 2] = res * (-float(center[1]) / h + .5)
    t[2, 2] = 1
    if not rot == 0:
        rot = -rot
        rot_mat = np.zeros((3, 3))
        rot_rad = rot * np.pi / 180
        sn, cs = np.sin(rot_rad), np.cos(rot_rad)
        rot_mat[0, :2] = [cs, -sn]
        rot_mat[1, :2] = [sn, cs]
        rot_mat[2, 2] = 1
        t_mat = np.eye(3)
        t_mat[0, 2] = -res / 2
        t_mat[1, 2] = -res / 2
        t_inv = t_mat.copy()
        t_inv[:2, 2] *= -1
        t = np.dot(t_inv, np.dot(rot_mat, np.dot(t_mat, t)))
    return t


def transform_pts(pts, center, scale, rot, res, size, invert=0):
    """pylib/Evaluation.py:179-187 -- Evaluation.py's OWN TransformPts (1-based points: ``pts - 1`` in,
    ``astype(int) + 1`` out), not the one of pylib/HumanAug.py:44-52 which lacks the two shifts."""
    nlmk = pts.shape[0]
    t = get_transform(center, scale, rot, res, size)
    if invert:
        t = np.linalg.inv(t)
    new_pt = np.concatenate((pts - 1, np.ones((nlmk, 1))), axis=1).T
    new_pt = np.dot(t, new_pt)
    return new_pt[0:2, :].T.astype(int) + 1


def final_preds(output, center, scale, res, rot):
    """pylib/Evaluation.py:108-132 + transform_preds :134-150: quarter-pixel refinement towards the higher
    neighbour, +0.5, inverse crop transform (size 200), truncation to int."""
    coords = get_preds(output)
    for n in range(coords.size(0)):
        for p in range(coords.size(1)):
            hm = output[n][p]
            px = int(math.floor(coords[n][p][0]))
            py = int(math.floor(coords[n][p][1]))
            if px > 1 and px < res[0] and py > 1 and py < res[1]:
                diff = torch.Tensor([hm[py - 1][px] - hm[py - 1][px - 2], hm[py][px - 1] - hm[py - 2][px - 1]])
                coords[n][p] += diff.sign() * .25
    coords += 0.5
    preds = coords.clone()
    for i in range(coords.size(0)):
        preds[i] = torch.from_numpy(transform_pts(coords[i].numpy(), center[i].numpy(), scale[i].numpy(),
                                                  rot[i].numpy(), res[0], size=200, invert=1))
    return preds


def accuracy_origin_res(output, center, scale, res, grnd_pts, normalizers, rot):
    """pylib/Evaluation.py:88-106 (PCKh at the original resolution, the 14 MPII joints of :92)."""
    idxs = [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15]
    pred_pts = final_preds(output, center, scale, res, rot)
    dists = calc_dists(pred_pts, grnd_pts, normalizers, use_zero=True)
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


def draw_gaussian(img, pt, sigma):
    """pylib/HumanPts.py:50-76 (numpy, in place)."""
    tmp_size = np.ceil(3 * sigma)
    ul = [int(pt[0] - tmp_size), int(pt[1] - tmp_size)]
    br = [int(pt[0] + tmp_size), int(pt[1] + tmp_size)]
    if (ul[0] >= img.shape[1] or ul[1] >= img.shape[0] or br[0] < 0 or br[1] < 0):
        return img
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    g = np.exp(- ((x - x0) ** 2 + (y - y0) ** 2) / (tmp_size ** 2))
    g_x = max(0, -ul[0]), min(br[0] + 1, img.shape[1]) - max(0, ul[0]) + max(0, -ul[0])
    g_y = max(0, -ul[1]), min(br[1] + 1, img.shape[0]) - max(0, ul[1]) + max(0, -ul[1])
    img_x = max(0, ul[0]), min(br[0] + 1, img.shape[1])
    img_y = max(0, ul[1]), min(br[1] + 1, img.shape[0])
    img[img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
    return img


def pts2heatmap(pts, heatmap_shape, sigma=1):
    """pylib/HumanPts.py:35-48."""
    heatmap = np.zeros((pts.shape[0], heatmap_shape[0], heatmap_shape[1]))
    valid_pts = np.zeros((pts.shape))
    for i in range(0, pts.shape[0]):
        if pts[i][0] <= 0 or pts[i][1] <= 0:
            continue
        heatmap[i] = draw_gaussian(heatmap[i], pts[i], sigma)
        valid_pts[i] = pts[i]
    return heatmap, valid_pts
