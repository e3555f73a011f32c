"""Drop-in for the two ``pylib/HumanAug.py`` functions on the validation path (flip test-time augmentation,
cu-net.py:240-249): ``flip_channels`` (:198-210) and ``shuffle_channels_for_horizontal_flipping`` (:177-196).

Device-agnostic tensor ops (the reference round-trips through numpy on the CPU every iteration): a width flip and one
channel gather.  Both return a new tensor (the reference's shuffle mutates its argument and returns it; every caller
uses the return value)."""
import torch

# cu-net.py:32-33 -- MPII left/right joint pairs swapped by a horizontal flip
MPII_FLIP_INDEX = ((1, 4), (0, 5), (12, 13), (11, 14), (10, 15), (2, 3))


def flip_channels(maps):
    """Horizontally flip every map: [N,C,H,W] or [C,H,W] (HumanAug.py:198-210)."""
    if maps.dim() not in (3, 4):
        raise ValueError("tensor dimension is not right")
    return torch.flip(maps, dims=[-1]).float()


def flip_permutation(num_channels, flip_indxs=MPII_FLIP_INDEX):
    """Channel permutation equivalent to the reference's sequence of pairwise swaps (HumanAug.py:188-192)."""
    perm = list(range(num_channels))
    for a, b in (tuple(int(v) for v in pair) for pair in flip_indxs):
        perm[a], perm[b] = perm[b], perm[a]
    return perm


def shuffle_channels_for_horizontal_flipping(maps, flip_indxs=MPII_FLIP_INDEX):
    """Swap the left/right channels of [N,C,H,W] or [C,H,W] maps (HumanAug.py:177-196)."""
    if maps.dim() == 4:
        dim = 1
    elif maps.dim() == 3:
        dim = 0
    else:
        raise ValueError("tensor dimension is not right")
    perm = torch.tensor(flip_permutation(maps.shape[dim], flip_indxs), device=maps.device, dtype=torch.long)
    return maps.index_select(dim, perm)
