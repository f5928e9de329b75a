"""MaskHeadSegLevel / MaskPredictionLayer (modules/heads/mask_head.py:10-57) on the sm_100a kernels.

Called from inside the decoder loop through a functools.partial (model/query3d_unified.py:176-180):
`mask_head(query) -> (cls_logits (B,N,C), mask_logits (B,S,N), attn_mask (B,N,S) bool)`.

What is hoisted: `k_proj_m(feat_m)` does not depend on the query, so it is computed once per set of
segment features (the reference recomputes it on each of the K*L+1 calls) into one concatenated
operand Kcat[b*S+s, m*D:(m+1)*D] with rows of invalid tokens zeroed — the sum over memories of
`logits_m * valid_m` then is a single K = n_mem*D contraction per scene.
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops

bf16 = torch.bfloat16


class MaskPredictionLayer(nn.Module):
    """modules/heads/mask_head.py:46-57 (parameter container)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.q_proj = nn.Linear(hidden_size, hidden_size)
        self.k_proj = nn.Linear(hidden_size, hidden_size, False)


def mlp_head_params(input_size, hidden_size, output_size, dropout=0.0) -> nn.Sequential:
    """Same module indices as get_mlp_head (modules/utils.py:18-25): 0 Linear, 2 LayerNorm(1e-12), 4 Linear."""
    return nn.Sequential(nn.Linear(input_size, hidden_size), nn.ReLU(), nn.LayerNorm(hidden_size, eps=1e-12),
                         nn.Dropout(dropout), nn.Linear(hidden_size, output_size))


class MlpHeadRunner:
    """Linear-ReLU-LayerNorm-Linear through the GEMM / LayerNorm kernels (eval: dropout is identity)."""

    def __init__(self, seq: nn.Sequential, neg_inf_cols=None):
        self.seq, self.neg_inf_cols = seq, neg_inf_cols
        self._key, self._w, self._ws = None, None, {}

    def _weights(self, dev):
        ps = list(self.seq.parameters())
        key = (str(dev), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if self._key != key:
            l0, ln, l4 = self.seq[0], self.seq[2], self.seq[4]
            b4 = l4.bias.detach().float().clone().to(dev)
            if self.neg_inf_cols is not None:
                b4[..., self.neg_inf_cols] = float("-inf")       # cls_logits[..., filter] = -inf, folded into the bias
            self._w = dict(w0=l0.weight.detach().to(dev, bf16).contiguous(), b0=l0.bias.detach().float().to(dev),
                           g=ln.weight.detach().float().to(dev)[None].contiguous(),
                           be=ln.bias.detach().float().to(dev)[None].contiguous(), eps=ln.eps,
                           w4=l4.weight.detach().to(dev, bf16).contiguous(), b4=b4.contiguous())
            self._key = key
        return self._w

    def __call__(self, x16: torch.Tensor, R: int) -> torch.Tensor:
        """x16: bf16 [R, Din] -> fp32 [R, Dout] (a fresh tensor)."""
        dev = x16.device
        w = self._weights(dev)
        Hd, Din, Dout = w["w0"].shape[0], w["w0"].shape[1], w["w4"].shape[0]
        h = torch.empty(R, Hd, dtype=torch.float32, device=dev)
        ops.linear(x16, w["w0"], h, M=R, N=Hd, K=Din, bias=w["b0"], relu=True)
        h16 = torch.empty(R, Hd, dtype=bf16, device=dev)
        ops.add_layernorm(h, None, w["g"], w["be"], w["eps"], R, Hd, out_bf16=h16)
        out = torch.empty(R, Dout, dtype=torch.float32, device=dev)
        ops.linear(h16, w["w4"], out, M=R, N=Dout, K=Hd, bias=w["b4"])
        return out


class MaskHeadSegLevel(nn.Module):
    """Drop-in for modules/heads/mask_head.py:10-44 (same kwargs, state_dict keys and forward)."""

    def __init__(self, cfg=None, hidden_size=768, num_targets=201, memories_for_match=["voxel"],
                 filter_out_classes=None, dropout=0.1):
        super().__init__()
        self.cls_head = mlp_head_params(hidden_size, hidden_size, num_targets, dropout=dropout)
        self.filter_out_classes = filter_out_classes
        memories_for_match = [m for m in memories_for_match if m in ("voxel", "mv", "pc")]
        layer = MaskPredictionLayer(hidden_size)
        n = len(memories_for_match)
        self.mask_pred_list = nn.ModuleList([copy.deepcopy(layer) for _ in range(n - 1)] + [layer])   # layer_repeat
        self.hidden_size = hidden_size
        # the reference indexes with `filter_out_classes` even when it is None, which then addresses
        # every class (`x[..., None] = -inf`); keep that behaviour
        cols = slice(None) if filter_out_classes is None else list(filter_out_classes)
        self._cls = MlpHeadRunner(self.cls_head, neg_inf_cols=cols)
        self._wkey, self._w = None, None
        self._kcat_key, self._kcat = None, None

    def _weights(self, dev):
        ps = [p for l in self.mask_pred_list for p in l.parameters()]
        key = (str(dev), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if self._wkey != key:
            self._w = dict(
                wq=torch.cat([l.q_proj.weight.detach() for l in self.mask_pred_list], 0).to(dev, bf16).contiguous(),
                bq=torch.cat([l.q_proj.bias.detach() for l in self.mask_pred_list], 0).float().to(dev).contiguous(),
                wk=[l.k_proj.weight.detach().to(dev, bf16).contiguous() for l in self.mask_pred_list])
            self._wkey, self._kcat_key = key, None
        return self._w

    def _segment_keys(self, seg_fts_for_match, w, dev):
        """Kcat [B*S, n*D] = concat_m valid_m * k_proj_m(feat_m), cached per set of feature tensors."""
        key = tuple((f.data_ptr(), f._version, m.data_ptr(), m._version) for f, m, _ in seg_fts_for_match)
        if self._kcat_key != key:
            n, D = len(seg_fts_for_match), self.hidden_size
            B, S = seg_fts_for_match[0][0].shape[:2]
            kcat = torch.empty(B * S, n * D, dtype=bf16, device=dev)
            x16 = torch.empty(B * S, D, dtype=bf16, device=dev)
            masks = []
            for j, (feat, mask, _pos) in enumerate(seg_fts_for_match):
                if mask.ndim != 2 or mask.dtype != torch.bool:
                    raise ValueError("mask head: per-memory masks must be bool (B, S), True = ignore")
                mask = mask.contiguous()
                masks.append(mask)
                ops.ingest_memory(feat.contiguous().float(), None, None, x16, S)
                ops.linear(x16, w["wk"][j], kcat[:, j * D:(j + 1) * D], M=B * S, N=D, K=D, ldc=n * D, row_zero=mask)
            ptrs = torch.tensor([m.data_ptr() for m in masks], dtype=torch.int64, device=dev)
            self._kcat, self._kcat_key = (kcat, masks, ptrs, B, S), key
        return self._kcat

    def forward(self, query, seg_fts_for_match, seg_masks, offline_attn_masks=None, skip_prediction=False):
        if skip_prediction:
            return None, None, offline_attn_masks
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("pq3d_b200.MaskHeadSegLevel: inference path only — call under torch.no_grad()")
        dev = query.device
        B, N, D = query.shape
        R = B * N
        n = len(self.mask_pred_list)
        w = self._weights(dev)
        kcat, masks, ptrs, Bk, S = self._segment_keys(list(seg_fts_for_match)[:n], w, dev)
        x16 = torch.empty(R, D, dtype=bf16, device=dev)
        ops.cast_bf16(query.reshape(R, D).contiguous().float(), x16)
        cls_logits = self._cls(x16, R).view(B, N, -1)
        qcat = torch.empty(R, n * D, dtype=bf16, device=dev)
        ops.linear(x16, w["wq"], qcat, M=R, N=n * D, K=D, bias=w["bq"])
        raw = torch.empty(B, S, N, dtype=torch.float32, device=dev)
        ops.linear(kcat, qcat, raw, M=S, N=N, K=n * D, groups=B, a_group_rows=S, w_group_rows=N, ldc=N,
                   c_group_stride=S * N)
        mask_logits = torch.empty(B, S, N, dtype=torch.float32, device=dev)
        attn_mask = torch.empty(B, N, S, dtype=torch.bool, device=dev)
        ops.mask_head_finalize(raw, ptrs, n, seg_masks.contiguous(), mask_logits, attn_mask, B, S, N)
        if offline_attn_masks is not None:
            attn_mask = offline_attn_masks
        return cls_logits, mask_logits, attn_mask
