"""Voxel -> segment pooling + per-scale Linear + LayerNorm: everything
`PCDMask3DSegLevelEncoder.forward` (modules/vision/pcd_mask3d_encoder.py:116-154) does AFTER its MinkowskiEngine
backbone — the step directly upstream of the decoder's multi-scale `voxel` memory.

    for hlevel, feat_proj in zip(self.hlevels, self.feat_proj_list):
        batch_feat = stack([scatter_mean(f, p2s, dim=0, dim_size=max_seg) for f, p2s in zip(decomposed, point2segment)])
        multi_scale_seg_feats.append(feat_proj(batch_feat))              # Linear(C_h -> hidden) + LayerNorm + Dropout

The sparse-conv backbone (Res16UNet34C) and the transposed pooling that brings every scale to full voxel resolution are
out of scope (MinkowskiEngine); this module takes their output — one fp32 (Nv_total, C_h) table per scale, scenes
concatenated the way MinkowskiEngine stores them — and keeps the reference's parameter names
(`feat_proj_list.{i}.{0,1}.{weight,bias}`), so that part of a reference checkpoint loads unchanged.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from . import ops
from .query3d_unified import LinearLN

bf16 = torch.bfloat16
PLANES_LAST5 = (256, 256, 128, 96, 96)      # Res16UNet34C.PLANES[-5:] (modules/third_party/mask3d/models: the `sizes` of :123)


class SegLevelPooling(nn.Module):
    def __init__(self, cfg=None, hidden_size: int = 768, hlevels: Sequence[int] = (0, 1, 2, 3), dropout: float = 0.1,
                 sizes: Sequence[int] = PLANES_LAST5):
        super().__init__()
        self.hlevels = list(hlevels) + [4]          # 4 = the last level, always used for the mask features (:124)
        self.sizes = list(sizes)
        self.feat_proj_list = nn.ModuleList()
        for h in self.hlevels:
            seq = LinearLN(self.sizes[h], hidden_size)
            seq.append(nn.Dropout(dropout))          # index 2, parameter-free: state_dict keys match the reference's
            self.feat_proj_list.append(seq)
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    m.weight.normal_(0.0, 0.02)
                    m.bias.zero_()

    def forward(self, feats: Sequence[torch.Tensor], voxel_offsets: Sequence[int], point2segment, max_seg: int) -> List[torch.Tensor]:
        """feats[i]: fp32 (Nv_total, sizes[hlevels[i]]) full-resolution voxel features of scale hlevels[i];
        point2segment: list of per-scene int64 (Nv_b,) tensors (as the reference passes them) or their concatenation.
        Returns the reference's `multi_scale_seg_feats`: one (B, max_seg, hidden) fp32 tensor per scale."""
        if isinstance(point2segment, (list, tuple)):
            point2segment = torch.cat([p.reshape(-1) for p in point2segment])
        p2s = point2segment.to(torch.int64).contiguous()
        B = len(voxel_offsets) - 1
        perm, offsets = ops.segment_csr(p2s, voxel_offsets, max_seg)          # index work once, shared by all scales
        outs = []
        for f, proj in zip(feats, self.feat_proj_list):
            if f.requires_grad and torch.is_grad_enabled():
                raise NotImplementedError("SegLevelPooling: gradients into the voxel backbone are out of scope (the "
                                          "MinkowskiEngine backbone is not part of this package); detach its features")
            w = proj._weights(f.device)
            x16 = torch.empty(B * max_seg, w["k"], dtype=bf16, device=f.device)
            ops.segment_mean(f.detach().float().contiguous(), perm, offsets, out16=x16)   # pooled rows land as the GEMM operand
            if proj._needs_grad():
                from .train_blocks import linear_ln_train
                y = linear_ln_train(None, proj[0], proj[1], x16=x16)
            else:
                y = proj.run16(x16, B * max_seg, torch.empty(B * max_seg, w["w"].shape[0], dtype=torch.float32, device=f.device))
            y = y.view(B, max_seg, -1)
            if self.training:
                y = proj[2](y)
            outs.append(y)
        return outs

    def pooled(self, feat: torch.Tensor, voxel_offsets: Sequence[int], point2segment: torch.Tensor, max_seg: int) -> torch.Tensor:
        """scatter_mean alone: fp32 (B, max_seg, C)."""
        perm, offsets = ops.segment_csr(point2segment.to(torch.int64).contiguous(), voxel_offsets, max_seg)
        out = torch.empty((len(voxel_offsets) - 1) * max_seg, feat.shape[1], dtype=torch.float32, device=feat.device)
        ops.segment_mean(feat.float().contiguous(), perm, offsets, out32=out)
        return out.view(len(voxel_offsets) - 1, max_seg, -1)
