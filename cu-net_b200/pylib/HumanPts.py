"""Drop-in for the landmark <-> heat-map conversions of ``pylib/HumanPts.py`` (pts2heatmap :35-48, draw_gaussian
:50-76, heatmap2pts :91-110) as batched tensor ops that run where their inputs live: the reference draws every target
heat map in a Python loop on the CPU inside the DataLoader, which cannot feed a B200.

``pts2heatmap(pts[..., C, 2], (H, W), sigma)`` -> ``(heatmaps[..., C, H, W], valid_pts)``: one un-normalised Gaussian
``exp(-(dx^2 + dy^2) / ceil(3 sigma)^2)`` per landmark (peak 1.0), window and centre computed exactly as the reference
does -- ``int()`` truncation towards zero of ``pt -+ ceil(3 sigma)``, so a landmark closer than the window radius to
the left/top border is drawn one pixel off, faithfully -- clipped to the map, nothing for landmarks with x <= 0 or
y <= 0.
"""
import math

import torch


def pts2heatmap(pts, heatmap_shape, sigma=1):
    pts = torch.as_tensor(pts)
    lead = pts.shape[:-1]
    p = pts.reshape(-1, 2).double()
    hh, ww = int(heatmap_shape[0]), int(heatmap_shape[1])
    t = float(math.ceil(3 * sigma))
    ul = torch.trunc(p - t)                          # int(pt - tmp_size): truncation towards zero (HumanPts.py:53)
    br = torch.trunc(p + t)
    present = (p[:, 0] > 0) & (p[:, 1] > 0)          # HumanPts.py:43; such a landmark is "valid" even when its window
    draw = present & ~((ul[:, 0] >= ww) | (ul[:, 1] >= hh) | (br[:, 0] < 0) | (br[:, 1] < 0))   # misses the map (:56-59)
    xs = torch.arange(ww, dtype=torch.float64, device=p.device).view(1, 1, ww)
    ys = torch.arange(hh, dtype=torch.float64, device=p.device).view(1, hh, 1)
    cx, cy = (ul[:, 0] + t).view(-1, 1, 1), (ul[:, 1] + t).view(-1, 1, 1)      # g is centred at index size // 2
    inside = (xs >= ul[:, 0].view(-1, 1, 1)) & (xs <= br[:, 0].view(-1, 1, 1)) & \
             (ys >= ul[:, 1].view(-1, 1, 1)) & (ys <= br[:, 1].view(-1, 1, 1)) & draw.view(-1, 1, 1)
    g = torch.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (t * t))
    hm = torch.where(inside, g, torch.zeros_like(g)).float()
    valid = torch.where(present.view(-1, 1), p, torch.zeros_like(p)).to(pts.dtype if pts.is_floating_point() else torch.float32)
    return hm.reshape(*lead, hh, ww), valid.reshape(*lead, 2)


def heatmap2pts(heatmap):
    """HumanPts.py:91-110: x = idx % W, y = floor(idx / W) + 0.5 (sic: only y gets the half pixel), zero where max <= 0."""
    b, n, h, w = heatmap.shape
    mx, idx = torch.max(heatmap.reshape(b, n, h * w), 2)
    pts = torch.zeros(b, n, 2, device=heatmap.device)
    pts[:, :, 0] = (idx % w).float()
    pts[:, :, 1] = torch.div(idx, w, rounding_mode="floor").float() + 0.5
    return pts * mx.gt(0).unsqueeze(-1).float()
