"""Matcher cost matrices and matched mask losses on the sm_100a kernels — the step directly downstream
of the decoder's `predictions_class` / `predictions_mask` in stage-1 training.

`HungarianMatcher` mirrors modules/third_party/mask3d/matcher.py:67-193 (same constructor, `forward(outputs, targets,
mask_type)` / `memory_efficient_forward`, same return value: one `(query_idx, target_idx)` int64 pair per scene).  The
reference assembles three N x M x S einsums per scene in fp32 and moves every cost matrix to the host one scene at a
time; here ONE kernel launch (`pq3d_match_cost`) produces the cost matrices of the whole batch, a single device->host
copy follows, and the assignment itself stays scipy's `linear_sum_assignment` on the host, exactly as in the reference
(matcher.py:184) — the LSAP is sequential and tiny (100 x ~30).

`matched_mask_losses` mirrors `SetCriterion.loss_masks` (criterion.py:163-196) for `num_points = -1`.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


def pack_targets(targets: Sequence[dict], mask_type: str, S: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """List of per-scene {labels (M_b,), mask_type: (M_b, S)} -> padded uint8 masks (B, Mmax, S), int64 labels (B, Mmax),
    int32 counts (B,)."""
    B = len(targets)
    Mmax = max(1, max(int(t["labels"].shape[0]) for t in targets))
    masks = torch.zeros(B, Mmax, S, dtype=torch.uint8, device=device)
    labels = torch.zeros(B, Mmax, dtype=torch.int64, device=device)
    counts = torch.zeros(B, dtype=torch.int32)
    for b, t in enumerate(targets):
        m = t[mask_type]
        M = int(t["labels"].shape[0])
        if M:
            if m.shape[1] != S:
                raise ValueError(f"target masks of scene {b} have {m.shape[1]} points, predictions {S}")
            masks[b, :M] = (m != 0).to(torch.uint8)
            labels[b, :M] = t["labels"].to(torch.int64)
        counts[b] = M
    return masks, labels, counts.to(device)


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, num_points: int = 0,
                 ignore_label: int = -100):
        super().__init__()
        self.cost_class, self.cost_mask, self.cost_dice = cost_class, cost_mask, cost_dice
        self.ignore_label = ignore_label
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points
        if num_points != -1:
            raise NotImplementedError("pq3d_b200.HungarianMatcher matches on all points (num_points = -1, the shipped "
                                      "configuration, configs/instseg_sceneverse.yaml:174); random point subsets are not built")

    @torch.no_grad()
    def cost_matrices(self, outputs, targets, mask_type) -> Tuple[torch.Tensor, List[int]]:
        """(B, N, Mmax) fp32 cost on the device + the number of targets per scene."""
        pred_masks = outputs["pred_masks"].detach().float().contiguous()      # (B, S, N)
        pred_logits = outputs["pred_logits"].detach().float().contiguous()    # (B, N, C)
        masks, labels, counts = pack_targets(targets, mask_type, pred_masks.shape[1], pred_masks.device)
        cost = ops.match_cost(pred_masks, pred_logits, masks, labels, counts, self.cost_class, self.cost_mask,
                              self.cost_dice, self.ignore_label)
        return cost, [int(t["labels"].shape[0]) for t in targets]

    @torch.no_grad()
    def memory_efficient_forward(self, outputs, targets, mask_type):
        from scipy.optimize import linear_sum_assignment
        cost, counts = self.cost_matrices(outputs, targets, mask_type)
        cost = cost.cpu()                                                     # one copy for the whole batch
        indices = []
        for b, M in enumerate(counts):
            i, j = linear_sum_assignment(cost[b, :, :M])
            indices.append((torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)))
        return indices

    @torch.no_grad()
    def forward(self, outputs, targets, mask_type):
        return self.memory_efficient_forward(outputs, targets, mask_type)


class _MatchedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_masks, tgt_masks, pairs):
        pm = pred_masks.detach().float().contiguous()
        ce, dice, sums = ops.matched_mask_loss_fwd(pm, tgt_masks, pairs)
        ctx.save_for_backward(pm, tgt_masks, pairs, sums)
        ctx.in_dtype = pred_masks.dtype
        return ce, dice

    @staticmethod
    def backward(ctx, g_ce, g_dice):
        pm, tgt_masks, pairs, sums = ctx.saved_tensors
        z = lambda g: torch.zeros(pairs.shape[0], dtype=torch.float32, device=pm.device) if g is None else g.float().contiguous()  # noqa: E731
        d = ops.matched_mask_loss_bwd(pm, tgt_masks, pairs, z(g_ce), z(g_dice), sums)
        return d.to(ctx.in_dtype), None, None


def matched_mask_losses(pred_masks: torch.Tensor, targets: Sequence[dict], indices, mask_type: str = "segment_masks"):
    """criterion.py:163-196: {"loss_mask", "loss_dice"} = mean over scenes of (sum over that scene's matched pairs /
    number of pairs).  pred_masks (B, S, N) — a `predictions_mask` entry; indices from the matcher."""
    B, S, N = pred_masks.shape
    dev = pred_masks.device
    masks, _, _ = pack_targets(targets, mask_type, S, dev)
    rows, weights = [], []
    for b, (qi, ti) in enumerate(indices):
        k = int(qi.numel())
        for q, t in zip(qi.tolist(), ti.tolist()):
            rows.append((b, q, t))
            weights.append(1.0 / (k * B))
    if not rows:
        zero = pred_masks.sum() * 0.0
        return {"loss_mask": zero, "loss_dice": zero}
    pairs = torch.tensor(rows, dtype=torch.int32, device=dev)
    w = torch.tensor(weights, dtype=torch.float32, device=dev)
    ce, dice = _MatchedLoss.apply(pred_masks, masks, pairs)
    return {"loss_mask": (ce * w).sum(), "loss_dice": (dice * w).sum()}
