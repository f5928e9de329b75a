"""Inference post-processing of the decoder's instance predictions on the sm_100a kernels: what
`InstSegEval.eval_instance_step` (evaluator/instseg_eval.py:85-150) does per scene between the model's output and the
metric accumulation — class softmax, top-k over (query, class), mask scores, full-resolution masks / heatmaps, ordering
by score — except the optional DBSCAN split (`use_dbscan`, scikit-learn on the host) and the dataset-specific label
remapping.  Nothing voxel-sized is materialised; see csrc/postprocess.cu for the data-flow argument."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib, ops


def _i32(n, dev):
    return torch.empty(n, dtype=torch.int32, device=dev)


def instseg_postprocess(pred_logits: torch.Tensor, pred_masks: torch.Tensor, voxel2segment: torch.Tensor,
                        voxel_to_full: torch.Tensor, segment_to_full: torch.Tensor, topk_per_scene: int = -1,
                        with_heatmap: bool = True) -> Dict[str, torch.Tensor]:
    """One scene.  pred_logits (Q, C+1) class logits (last = no object), pred_masks (S, Q) segment mask logits
    (`predictions_class[-1][b]`, `predictions_mask[-1][b]`), voxel2segment (V,), voxel_to_full (P,), segment_to_full (P,)
    int64.  Returns scores (K,) descending, classes (K,), masks (P, K) float 0/1, heatmap (P, K), query (K,)."""
    for t, nm in ((pred_logits, "pred_logits"), (pred_masks, "pred_masks")):
        ops._chk(t, torch.float32, nm, 2)
    for t, nm in ((voxel2segment, "voxel2segment"), (voxel_to_full, "voxel_to_full"), (segment_to_full, "segment_to_full")):
        ops._chk(t, torch.int64, nm, 1)
    pred_logits, pred_masks = pred_logits.contiguous(), pred_masks.contiguous()
    dev = pred_logits.device
    lib, st = _lib.lib(), ops._stream()
    Q, C1 = pred_logits.shape
    S = pred_masks.shape[0]
    C = C1 - 1
    K = Q if topk_per_scene == -1 else int(topk_per_scene)           # :286-289
    if K > 1024 or K > Q * C:
        raise NotImplementedError(f"top-k of {K} > 1024 candidates is not built")
    P = voxel_to_full.numel()
    probs = torch.empty(Q, C, dtype=torch.float32, device=dev)
    _lib.check(lib.pq3d_class_probs(pred_logits.data_ptr(), probs.data_ptr(), Q, C1, st), "pq3d_class_probs")
    cls_score, flat = torch.empty(K, dtype=torch.float32, device=dev), _i32(K, dev)
    _lib.check(lib.pq3d_topk(probs.data_ptr(), Q * C, K, cls_score.data_ptr(), flat.data_ptr(), st), "pq3d_topk")
    seg_count = _i32(S, dev)
    _lib.check(lib.pq3d_bincount(voxel2segment.contiguous().data_ptr(), voxel2segment.numel(), seg_count.data_ptr(), S, st),
               "pq3d_bincount")
    score = torch.empty(K, dtype=torch.float32, device=dev)
    _lib.check(lib.pq3d_instseg_scores(pred_masks.data_ptr(), seg_count.data_ptr(), flat.data_ptr(), C, S, Q, K,
                                       cls_score.data_ptr(), score.data_ptr(), st), "pq3d_instseg_scores")
    # final order by score (descending), known before the full-resolution pass
    sorted_score, order = torch.empty(K, dtype=torch.float32, device=dev), _i32(K, dev)
    _lib.check(lib.pq3d_topk(score.data_ptr(), K, K, sorted_score.data_ptr(), order.data_ptr(), st), "pq3d_topk")
    q_of, cls_of = _i32(K, dev), _i32(K, dev)
    _lib.check(lib.pq3d_split_index(flat.data_ptr(), order.data_ptr(), K, C, q_of.data_ptr(), cls_of.data_ptr(), st),
               "pq3d_split_index")
    n_full = int(segment_to_full.max().item()) + 1                    # scatter_mean's dim_size (torch_scatter default)
    points = _i32(n_full, dev)
    _lib.check(lib.pq3d_bincount(segment_to_full.contiguous().data_ptr(), P, points.data_ptr(), n_full, st), "pq3d_bincount")
    votes = _i32(n_full * K, dev)
    masks = torch.empty(P, K, dtype=torch.float32, device=dev)
    heat = torch.empty(P, K, dtype=torch.float32, device=dev) if with_heatmap else None
    _lib.check(lib.pq3d_instseg_fullres(pred_masks.data_ptr(), q_of.data_ptr(), voxel2segment.data_ptr(),
                                        voxel_to_full.contiguous().data_ptr(), segment_to_full.data_ptr(), P, S, Q, K, n_full,
                                        votes.data_ptr(), points.data_ptr(), masks.data_ptr(),
                                        None if heat is None else heat.data_ptr(), st), "pq3d_instseg_fullres")
    ops._count(9)
    return {"scores": sorted_score, "classes": cls_of.long(), "masks": masks, "heatmap": heat, "query": q_of.long()}
