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
