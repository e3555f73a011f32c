"""Seeded synthetic batches for `cu-net.py` without a dataset and for `bench.py`.

The oracle keeps its own copy (oracle/synthetic.py); tests/test_oracle_golden.py checks the two stay identical.

img     = U[0,1) float32 [N,3,256,256]   (reference images are float in [0,1], utils/imutils.py:31-36)
heatmap = zeros [N,C,64,64] with one un-normalised Gaussian exp(-(dx^2+dy^2)/9), 7x7 support,
          peak 1.0 per channel at a seeded integer centre in [4,60)^2
          (the reference's sigma=1 draw divides by tmp_size^2 = 9: pylib/HumanPts.py:49-76).
"""
import torch


def make_inputs(n, class_num, seed=0, res_in=256, res_out=64):
    gen = torch.Generator().manual_seed(seed)
    img = torch.rand(n, 3, res_in, res_in, generator=gen)
    cx = torch.randint(4, res_out - 4, (n, class_num), generator=gen)
    cy = torch.randint(4, res_out - 4, (n, class_num), generator=gen)
    r = torch.arange(-3, 4, dtype=torch.float32)
    g = torch.exp(-(r[None, :] ** 2 + r[:, None] ** 2) / 9.0)
    hm = torch.zeros(n, class_num, res_out, res_out)
    for i in range(n):
        for c in range(class_num):
            x, y = int(cx[i, c]), int(cy[i, c])
            hm[i, c, y - 3:y + 4, x - 3:x + 4] = g
    return img, hm
