"""CPU: the oracle restatements of the §8f rows (voxel -> segment pooling, matcher cost matrices, matched mask losses)
pinned against the LIVE reference where it is importable (modules/third_party/mask3d/{matcher,criterion}.py) and against
first principles where the arithmetic lives in an absent third-party dependency (torch_scatter)."""
import importlib

import pytest
import torch

from oracle import ref_loader, restatement as O

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


def _scene(g, N=20, S=70, M=6, C=11):
    pred_logits = torch.randn(N, C, generator=g) * 2
    pred_masks = torch.randn(S, N, generator=g) * 3
    pred_masks[S - 5:] = -1e6                                   # padded segments, as the mask head writes them
    tgt = (torch.rand(M, S, generator=g) < 0.2)
    tgt[:, S - 5:] = False
    labels = torch.randint(0, C - 1, (M,), generator=g)
    labels[1] = -100
    return pred_logits, pred_masks, labels, tgt


def test_scatter_mean_restatement_first_principles():
    g = torch.Generator().manual_seed(0)
    src = torch.randn(500, 12, generator=g)
    idx = torch.randint(0, 40, (500,), generator=g)
    idx[idx == 7] = 8                                            # an empty segment
    out = O.scatter_mean(src, idx, 48)
    for s in range(48):
        rows = src[idx == s]
        want = torch.zeros(12) if rows.shape[0] == 0 else rows.sum(0) / rows.shape[0]
        assert torch.allclose(out[s], want, atol=1e-6)
    assert (out[7] == 0).all() and (out[40:] == 0).all()
    # ascending-order fp32 summation, bit for bit
    acc = torch.zeros(48, 12)
    for i in range(500):
        acc[idx[i]] += src[i]
    cnt = torch.bincount(idx, minlength=48).clamp(min=1).float()
    assert torch.equal(out, acc / cnt[:, None])


@needs_ref
def test_matcher_cost_and_assignment_match_live_reference():
    ref_loader.load()
    rm = importlib.import_module("modules.third_party.mask3d.matcher")
    g = torch.Generator().manual_seed(1)
    scenes = [_scene(g, M=m) for m in (6, 3, 9)]
    w = dict(cost_class=2.0, cost_mask=5.0, cost_dice=2.0)
    matcher = rm.HungarianMatcher(num_points=-1, ignore_label=-100, **w)
    outputs = {"pred_logits": torch.stack([s[0] for s in scenes]), "pred_masks": torch.stack([s[1] for s in scenes])}
    targets = [{"labels": s[2], "segment_masks": s[3]} for s in scenes]
    ref_idx = matcher(outputs, targets, "segment_masks")
    from scipy.optimize import linear_sum_assignment
    for b, s in enumerate(scenes):
        cost = O.matcher_cost(s[0], s[1], s[2], s[3], **w)
        # the reference's own cost pieces
        om, tm = s[1].T.float(), s[3].float()
        assert torch.allclose(O.batch_dice_cost(om, tm), rm.batch_dice_loss(om, tm), atol=1e-6)
        assert torch.allclose(O.batch_sigmoid_ce_cost(om, tm), rm.batch_sigmoid_ce_loss(om, tm), rtol=1e-6, atol=1e-4)
        i, j = linear_sum_assignment(cost)
        assert torch.equal(torch.as_tensor(i), ref_idx[b][0]) and torch.equal(torch.as_tensor(j), ref_idx[b][1])


@needs_ref
def test_matched_mask_losses_match_live_reference():
    ref_loader.load()
    try:
        rc = importlib.import_module("modules.third_party.mask3d.criterion")
    except Exception as e:                                          # torchvision missing etc.
        pytest.skip(f"reference criterion not importable here: {e}")
    g = torch.Generator().manual_seed(2)
    scenes = [_scene(g, M=m) for m in (6, 4)]
    pred_masks = torch.stack([s[1] for s in scenes]).requires_grad_(True)
    targets = [{"labels": s[2], "segment_masks": s[3]} for s in scenes]
    indices = [(torch.tensor([3, 0, 7, 11]), torch.tensor([1, 0, 5, 2])), (torch.tensor([2, 9]), torch.tensor([3, 0]))]
    crit = rc.SetCriterion.__new__(rc.SetCriterion)
    crit.num_points = -1
    want = rc.SetCriterion.loss_masks(crit, {"pred_masks": pred_masks}, targets, indices, 1.0, "segment_masks")
    got = O.matched_mask_losses(pred_masks, [s[3] for s in scenes], indices)
    for k in ("loss_mask", "loss_dice"):
        assert torch.allclose(got[k], want[k], rtol=1e-6, atol=1e-7), k
