"""Host-side pieces of the validation path (cu-net.py:240-249, pylib/Evaluation.py:25-85, pylib/HumanAug.py:177-210)
against the loop-for-loop restatements in oracle/evaluation_oracle.py.  No GPU: only the device-agnostic arithmetic."""
import torch

from oracle import evaluation_oracle, synthetic


def test_flip_helpers_match_reference_semantics():
    from cunet_b200.pylib import HumanAug
    g = torch.Generator().manual_seed(0)
    for shape in [(2, 16, 8, 8), (16, 8, 8)]:
        maps = torch.randn(*shape, generator=g)
        assert torch.equal(HumanAug.flip_channels(maps), evaluation_oracle.flip_channels(maps.clone()))
        ref = evaluation_oracle.shuffle_channels_for_horizontal_flipping(maps.clone(), HumanAug.MPII_FLIP_INDEX)
        assert torch.equal(HumanAug.shuffle_channels_for_horizontal_flipping(maps), ref)
    # the swap is an involution
    perm = HumanAug.flip_permutation(16)
    assert [perm[p] for p in perm] == list(range(16))


def test_accuracy_arithmetic_matches_reference_loops():
    from cunet_b200.pylib import Evaluation
    idxs = [0, 1, 2, 3, 4, 5, 10, 11, 14, 15]                      # cu-net.py:127
    _, target = synthetic.make_inputs(3, 16, seed=5)
    g = torch.Generator().manual_seed(1)
    output = target + 0.4 * torch.randn(target.shape, generator=g)  # noisy prediction: some joints move
    output[0, 3] = -1.0                                             # a channel without positive maximum -> pred (0, 0)
    target[1, 4] = 0.0                                              # a missing ground-truth joint -> dist -1
    want = evaluation_oracle.accuracy(output, target, idxs)
    preds, gts = evaluation_oracle.get_preds(output), evaluation_oracle.get_preds(target)
    got = Evaluation.accuracy_from_preds(preds, gts, output.shape[3], idxs)
    assert torch.allclose(got, want, atol=1e-6), (got, want)
    d_ref = evaluation_oracle.calc_dists(preds, gts, torch.ones(3) * 6.4)
    assert torch.allclose(Evaluation.calc_dists(preds, gts, torch.ones(3) * 6.4), d_ref, atol=1e-6)
    assert (d_ref == -1).any() and (d_ref > 0).any()
    # all-invalid joint
    assert float(Evaluation.dist_acc(torch.full((4,), -1.0))) == -1.0


def test_final_preds_arithmetic_matches_reference_loops():
    """Quarter-pixel refinement + inverse crop transform (with and without rotation) against the reference's loops."""
    from cunet_b200.pylib import Evaluation
    n, c = 4, 16
    _, target = synthetic.make_inputs(n, c, seed=9)
    g = torch.Generator().manual_seed(2)
    output = target + 0.05 * torch.randn(target.shape, generator=g)
    output[1, 2] = 0.0
    output[1, 2, 0, 5] = 1.0                                           # a peak on the border: no refinement
    center = torch.rand(n, 2, generator=g) * 400 + 300
    scale = torch.rand(n, generator=g) * 2 + 0.8
    rot = torch.tensor([0.0, 25.0, -40.0, 0.0])
    res = [64, 64]
    want = evaluation_oracle.final_preds(output.clone(), center, scale, res, rot)
    got = Evaluation.final_preds_from_coords(output, evaluation_oracle.get_preds(output), center, scale, res, rot)
    # truncation to int: allow a 1-pixel difference only where the float result sits within 1e-6 of an integer
    assert (got - want).abs().max() <= 1.0
    assert ((got - want).abs() > 0).float().mean() < 0.01
    # PCKh arithmetic on top of it
    grnd = want + torch.randint(-6, 7, want.shape, generator=g).float()
    grnd[0, 3] = 0.0                                                   # missing joint
    norm = torch.rand(n, generator=g) * 20 + 40
    d_ref = evaluation_oracle.calc_dists(want, grnd, norm, use_zero=True)
    assert torch.allclose(Evaluation.calc_dists(want, grnd, norm, use_zero=True), d_ref, atol=1e-6)


def test_pts2heatmap_matches_reference_drawing():
    """Batched target heat maps == the reference's per-landmark numpy drawing, including clipped windows, the
    truncation quirk next to the top/left border, landmarks outside the map and missing (<= 0) landmarks."""
    import numpy as np
    from cunet_b200.pylib import HumanPts
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(3, 16, 2, generator=g) * 70 - 3            # some outside [0, 64), some negative
    pts[0, 0] = torch.tensor([2.5, 30.2])                        # pt - 3 in (-1, 0): int() truncates to 0
    pts[0, 1] = torch.tensor([63.9, 0.4])
    pts[0, 2] = torch.tensor([10.0, 12.0])                       # integer landmark (the synthetic inputs)
    pts[0, 3] = torch.tensor([66.2, 20.0])                       # window partly inside from the right
    pts[0, 4] = torch.tensor([80.0, 20.0])                       # entirely outside
    got, valid = HumanPts.pts2heatmap(pts, (64, 64), sigma=1)
    for n in range(3):
        want, wvalid = evaluation_oracle.pts2heatmap(pts[n].numpy().astype(np.float64), (64, 64), 1)
        assert np.abs(got[n].numpy() - want).max() < 1e-6, n
        assert np.abs(valid[n].numpy() - wvalid).max() < 1e-6
    # the synthetic heat maps of the bench are this drawing at integer landmarks
    assert float(got[0, 2].max()) == 1.0 and int(got[0, 2].argmax()) == 12 * 64 + 10
    # and heatmap2pts inverts it up to the reference's (x, y + 0.5) convention
    back = HumanPts.heatmap2pts(got[:1, 2:3])
    assert back.tolist() == [[[10.0, 12.5]]]
