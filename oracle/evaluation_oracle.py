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
