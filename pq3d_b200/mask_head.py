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
        self._prep = None
        self._prep_slots = {}
        self._bufs = {}

    def _weights(self, dev):
        ps = [p for l in self.mask_pred_list for p in l.parameters()]
        key = (str(dev), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if self._wkey != key:
            self._w = dict(
                wq=torch.cat([l.q_proj.weight.detach() for l in self.mask_pred_list], 0).to(dev, bf16).contiguous(),
                bq=torch.cat([l.q_proj.bias.detach() for l in self.mask_pred_list], 0).float().to(dev).contiguous(),
                wk=[l.k_proj.weight.detach().to(dev, bf16).contiguous() for l in self.mask_pred_list])
            self._wkey = key
        return self._w

    def prepare(self, seg_fts_for_match, seg_masks, slot=0):
        """Hoisted, query-independent part, once per forward (the reference redoes it on each of its K*L+1 calls):
        Kcat [B*S, n*D] = concat_m valid_m * k_proj_m(feat_m).  Everything lands in persistent buffers (static
        addresses for CUDA-graph capture); nothing is cached across calls — tensor identity is not a safe key."""
        feats = list(seg_fts_for_match)[:len(self.mask_pred_list)]
        dev = feats[0][0].device
        w = self._weights(dev)
        n, D = len(feats), self.hidden_size
        B, S = feats[0][0].shape[:2]
        key = ("prep", B, S, slot)         # slot: one workspace per caller context (decoder workspace / CUDA stream)
        kcat = self._ws(key, "kcat", (B * S, n * D), bf16, dev)
        x16 = self._ws(key, "x16", (B * S, D), bf16, dev)
        masks = self._ws(key, "masks", (n + 1, B, S), torch.bool, dev)
        for j, (feat, mask, _pos) in enumerate(feats):
            if mask.ndim != 2 or mask.dtype != torch.bool:
                raise ValueError("mask head: per-memory masks must be bool (B, S), True = ignore")
            masks[j].copy_(mask)
            ops.ingest_memory(feat.contiguous().float(), None, None, x16, S)
            ops.linear(x16, w["wk"][j], kcat[:, j * D:(j + 1) * D], M=B * S, N=D, K=D, ldc=n * D, row_zero=masks[j])
        if seg_masks.dtype != torch.bool:
            raise TypeError("mask head: seg_masks must be torch.bool (True = padded segment)")
        masks[n].copy_(seg_masks)
        ptrs = self._ws(key, "ptrs", (n,), torch.int64, dev)
        if not self._bufs[key].get("ptrs_set"):
            ptrs.copy_(torch.tensor([masks[j].data_ptr() for j in range(n)], dtype=torch.int64))
            self._bufs[key]["ptrs_set"] = True
        self._prep = (kcat, masks, ptrs, B, S)
        self._prep_slots[slot] = self._prep
        return self._prep

    def _ws(self, key, name, shape, dtype, dev):
        ws = self._bufs.setdefault(key, {})
        t = ws.get(name)
        if t is None:
            t = ws[name] = torch.empty(shape, dtype=dtype, device=dev)
        return t

    def run_into(self, q2d: torch.Tensor, B: int, N: int, out_cls: torch.Tensor, out_logits: torch.Tensor,
                 out_attn: torch.Tensor, slot=0):
        """The per-call part, after prepare(): reads and writes static addresses only (caller-provided outputs, a
        persistent workspace), so the decoder can capture it into its CUDA graph.  q2d: fp32 [B*N, D]."""
        dev = q2d.device
        D, R, n = self.hidden_size, B * N, len(self.mask_pred_list)
        w = self._weights(dev)
        kcat, masks, ptrs, Bk, S = self._prep_slots[slot]
        key = (B, N, S, slot)
        x16 = self._ws(key, "x16", (R, D), bf16, dev)
        ops.cast_bf16(q2d, x16)
        cw = self._cls._weights(dev)
        Hd = cw["w0"].shape[0]
        h = self._ws(key, "h", (R, Hd), torch.float32, dev)
        ops.linear(x16, cw["w0"], h, M=R, N=Hd, K=D, bias=cw["b0"], relu=True)
        h16 = self._ws(key, "h16", (R, Hd), bf16, dev)
        ops.add_layernorm(h, None, cw["g"], cw["be"], cw["eps"], R, Hd, out_bf16=h16)
        C = cw["w4"].shape[0]
        ops.linear(h16, cw["w4"], out_cls, M=R, N=C, K=Hd, bias=cw["b4"], ldc=C)
        qcat = self._ws(key, "qcat", (R, n * D), bf16, dev)
        ops.linear(x16, w["wq"], qcat, M=R, N=n * D, K=D, bias=w["bq"])
        raw = self._ws(key, "raw", (B, S, N), torch.float32, dev)
        ops.linear(kcat, qcat, raw, M=S, N=N, K=n * D, groups=B, a_group_rows=S, w_group_rows=N, ldc=N,
                   c_group_stride=S * N)
        ops.mask_head_finalize(raw, ptrs, n, masks[n], out_logits, out_attn, B, S, N, masks=masks)

    def forward(self, query, seg_fts_for_match, seg_masks, offline_attn_masks=None, skip_prediction=False):
        if skip_prediction:
            return None, None, offline_attn_masks
        if self.training or (torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                          or query.requires_grad)):
            # training: forward + backward composed from the same kernels (MaskHeadTrain); no autograd fallback
            feats = list(seg_fts_for_match)[:len(self.mask_pred_list)]
            meta = dict(feats=feats, seg_masks=seg_masks)
            cls_logits, mask_logits = _MaskHeadFunction.apply(self, meta, query, *[f[0] for f in feats],
                                                              *list(self.parameters()))
            attn_mask = meta["attn"] if offline_attn_masks is None else offline_attn_masks
            return cls_logits, mask_logits, attn_mask
        dev = query.device
        B, N, D = query.shape
        S = seg_fts_for_match[0][0].shape[1]
        C = self.cls_head[4].out_features
        cls_logits = torch.empty(B, N, C, dtype=torch.float32, device=dev)
        mask_logits = torch.empty(B, S, N, dtype=torch.float32, device=dev)
        attn_mask = torch.empty(B, N, S, dtype=torch.bool, device=dev)
        slot = ("stream", torch.cuda.current_stream(dev).cuda_stream)
        self.prepare(seg_fts_for_match, seg_masks, slot)
        self.run_into(query.reshape(B * N, D).contiguous().float(), B, N, cls_logits, mask_logits, attn_mask, slot)
        if offline_attn_masks is not None:
            attn_mask = offline_attn_masks
        return cls_logits, mask_logits, attn_mask


# ------------------------------------------------------------------------------------------------------------
# training: forward that keeps what the backward needs + the backward, composed from the same kernels
# ------------------------------------------------------------------------------------------------------------
def _pad64(n):
    return (n + 63) // 64 * 64


class MaskHeadTrain:
    """One forward pass' worth of mask-head calls in training (the decoder calls the head once per layer application,
    Query3DUnified once more at the end — modules/grounding/query_encoder.py:78-81, model/query3d_unified.py:184-190).

    The k-projection is hoisted as in inference; its gradient is the SUM over all calls of `d_raw q`, accumulated in
    fp32 and turned into `k_proj.weight` / feature gradients once, in `finish()`.  Everything GEMM-shaped runs on
    pq3d_linear_bf16 (dgrad with transposed weights, wgrad on K-major transposes); the cls head's Dropout
    (`MaskHeadSegLevel(dropout=0.1)`) uses the counter RNG (stream rng.SITE_MASK_HEAD + call index)."""

    def __init__(self, mh: "MaskHeadSegLevel", seg_fts_for_match, seg_masks, B, N, seed=None):
        from . import rng
        self.rng = rng
        self.mh, self.B, self.N = mh, B, N
        feats = list(seg_fts_for_match)[:len(mh.mask_pred_list)]
        self.n, self.D = len(feats), mh.hidden_size
        n, D = self.n, self.D
        dev = feats[0][0].device
        self.dev = dev
        if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
            raise NotImplementedError(
                "pq3d_b200: the mask head's training path builds host-side tables (mask pointer table, class-filter "
                "index) per forward and cannot be captured into a CUDA graph yet — run stage-1 training steps eagerly "
                "(GraphedTrainStep covers the decoder-only training step)")
        self.S = S = feats[0][0].shape[1]
        self.seed = seed
        from .train_blocks import MlpHeadTrain
        f = lambda t: t.detach().float().contiguous()                 # noqa: E731
        w16 = lambda t: t.detach().to(bf16).contiguous()              # noqa: E731
        cols = slice(None) if mh.filter_out_classes is None else list(mh.filter_out_classes)
        self.cls = MlpHeadTrain(mh.cls_head, cols, seed)
        self.C = self.cls.C
        self.w = dict(wq=w16(torch.cat([l.q_proj.weight for l in mh.mask_pred_list], 0)),
                      wqt=w16(torch.cat([l.q_proj.weight for l in mh.mask_pred_list], 0).t()),
                      bq=f(torch.cat([l.q_proj.bias for l in mh.mask_pred_list], 0)),
                      wk=[w16(l.k_proj.weight) for l in mh.mask_pred_list],
                      wkt=[w16(l.k_proj.weight.t()) for l in mh.mask_pred_list])
        self.masks = torch.empty(n + 1, B, S, dtype=torch.bool, device=dev)
        self.x16 = []
        self.kcat = torch.empty(B * S, n * D, dtype=bf16, device=dev)
        for j, (feat, mask, _pos) in enumerate(feats):
            if mask.ndim != 2 or mask.dtype != torch.bool:
                raise ValueError("mask head: per-memory masks must be bool (B, S), True = ignore")
            self.masks[j].copy_(mask)
            x = torch.empty(B * S, D, dtype=bf16, device=dev)
            ops.ingest_memory(feat.detach().contiguous().float(), None, None, x, S)
            ops.linear(x, self.w["wk"][j], self.kcat[:, j * D:(j + 1) * D], M=B * S, N=D, K=D, ldc=n * D,
                       row_zero=self.masks[j])
            self.x16.append(x)
        if seg_masks.dtype != torch.bool:
            raise TypeError("mask head: seg_masks must be torch.bool (True = padded segment)")
        self.masks[n].copy_(seg_masks)
        self.ptrs = torch.tensor([self.masks[j].data_ptr() for j in range(n)], dtype=torch.int64).to(dev)
        self.kcatT = None
        self.d_kcat = None
        self.grads = {}
        self.calls = 0

    def _acc(self, name, g):
        self.grads[name] = g if name not in self.grads else self.grads[name] + g

    def _site(self, call):
        return self.rng.SITE_MASK_HEAD + call

    # ---- one call -------------------------------------------------------------------------------------------
    def call(self, q2d: torch.Tensor):
        """q2d fp32 [B*N, D] -> (cls (B,N,C), mask_logits (B,S,N), attn_mask (B,N,S) bool, saved)."""
        B, N, S, n, D, dev, w = self.B, self.N, self.S, self.n, self.D, self.dev, self.w
        R = B * N
        x16 = torch.empty(R, D, dtype=bf16, device=dev)
        ops.cast_bf16(q2d, x16)
        call = self.calls
        self.calls += 1
        cls2d, csv = self.cls.fwd(x16, R, self._site(call))
        cls = cls2d.view(B, N, self.C)
        qcat = torch.empty(R, n * D, dtype=bf16, device=dev)
        ops.linear(x16, w["wq"], qcat, M=R, N=n * D, K=D, bias=w["bq"])
        raw = torch.empty(B, S, N, dtype=torch.float32, device=dev)
        ops.linear(self.kcat, qcat, raw, M=S, N=N, K=n * D, groups=B, a_group_rows=S, w_group_rows=N, ldc=N,
                   c_group_stride=S * N)
        logits = torch.empty(B, S, N, dtype=torch.float32, device=dev)
        attn = torch.empty(B, N, S, dtype=torch.bool, device=dev)
        ops.mask_head_finalize(raw, self.ptrs, n, self.masks[n], logits, attn, B, S, N, masks=self.masks)
        return cls, logits, attn, dict(x16=x16, cls=csv, qcat=qcat, call=call)

    # ---- its backward ---------------------------------------------------------------------------------------
    def _tcast(self, x, rows, cols, want_c=False, gate=None):
        xt = torch.empty(cols, _pad64(rows), dtype=bf16, device=self.dev)
        xc = torch.empty(rows, cols, dtype=bf16, device=self.dev) if want_c else None
        ops.transpose_cast(x, xt, xc, gate=gate)
        return xt, xc

    def _wgrad(self, dyT, xT, n_out, n_in):
        dW = torch.empty(n_out, n_in, dtype=torch.float32, device=self.dev)
        ops.linear(dyT, xT, dW, M=n_out, N=n_in, K=dyT.shape[1])
        return dW

    def _colsum(self, x, gate=None):
        out = torch.empty(x.shape[1], dtype=torch.float32, device=self.dev)
        ops.colsum(x, out, gate=gate)
        return out

    def call_bwd(self, sv, d_cls, d_logits):
        """-> d_q fp32 [B*N, D] (or None when neither prediction carries a gradient)."""
        B, N, S, n, D, dev, w = self.B, self.N, self.S, self.n, self.D, self.dev, self.w
        R = B * N
        parts = []
        xT = None
        if d_cls is not None:
            xT, _ = self._tcast(sv["x16"], R, D)
            parts.append(self.cls.bwd(sv["cls"], d_cls.detach().reshape(R, self.C).float(), xT))
        if d_logits is not None:
            Np, Sp = _pad64(N), _pad64(S)
            d_raw16 = torch.empty(B * S, Np, dtype=bf16, device=dev)
            ops.mask_head_finalize_bwd(d_logits.detach().float().contiguous(), self.masks, n, d_raw16, B, S, N)
            d_rawT = torch.empty(B, 1, Np, Sp, dtype=bf16, device=dev)
            ops.transpose_cast(d_raw16.view(B, 1, S, Np), d_rawT)
            if self.kcatT is None:
                self.kcatT = torch.empty(B, 1, n * D, Sp, dtype=bf16, device=dev)
                ops.transpose_cast(self.kcat.view(B, 1, S, n * D), self.kcatT)
            d_qcat = torch.empty(R, n * D, dtype=torch.float32, device=dev)
            ops.linear(d_rawT.view(B * Np, Sp), self.kcatT.view(B * n * D, Sp), d_qcat, M=N, N=n * D, K=Sp, groups=B,
                       a_group_rows=Np, w_group_rows=n * D, ldc=n * D, c_group_stride=N * n * D)
            qcatT = torch.empty(B, 1, n * D, Np, dtype=bf16, device=dev)
            ops.transpose_cast(sv["qcat"].view(B, 1, N, n * D), qcatT)
            tmp = torch.empty(B * S, n * D, dtype=torch.float32, device=dev)
            q2 = qcatT.view(B * n * D, Np)
            for j in range(n):
                ops.linear(d_raw16, q2[j * D:], tmp[:, j * D:(j + 1) * D], M=S, N=D, K=Np, groups=B, a_group_rows=S,
                           w_group_rows=n * D, ldc=n * D, c_group_stride=S * n * D, row_zero=self.masks[j],
                           row_zero_group_stride=S)
            if self.d_kcat is None:
                self.d_kcat = tmp
            else:
                ops.add3(self.d_kcat, tmp, None, self.d_kcat)
            d_bq = self._colsum(d_qcat)
            d_qcatT, d_qcat16 = self._tcast(d_qcat, R, n * D, want_c=True)
            if xT is None:
                xT, _ = self._tcast(sv["x16"], R, D)
            d_wq = self._wgrad(d_qcatT, xT, n * D, D)
            for j in range(n):
                self._acc(f"mask_pred_list.{j}.q_proj.weight", d_wq[j * D:(j + 1) * D])
                self._acc(f"mask_pred_list.{j}.q_proj.bias", d_bq[j * D:(j + 1) * D])
            d_xb = torch.empty(R, D, dtype=torch.float32, device=dev)
            ops.linear(d_qcat16, w["wqt"], d_xb, M=R, N=D, K=n * D)
            parts.append(d_xb)
        if not parts:
            return None
        if len(parts) == 1:
            return parts[0]
        d_q = torch.empty(R, D, dtype=torch.float32, device=dev)
        ops.add3(parts[0], parts[1], None, d_q)
        return d_q

    def finish(self, want_feat_grads=True):
        """k_proj weight gradients (and d_feat per memory) from the accumulated d_kcat; returns (param grads, d_feats)."""
        B, S, n, D, dev, w = self.B, self.S, self.n, self.D, self.dev, self.w
        d_feats = [None] * n
        if self.d_kcat is not None:
            rows = B * S
            for j in range(n):
                dk = self.d_kcat[:, j * D:(j + 1) * D]
                dkT = torch.empty(D, _pad64(rows), dtype=bf16, device=dev)
                dk16 = torch.empty(rows, D, dtype=bf16, device=dev)
                ops.transpose_cast(dk, dkT, dk16)
                xT, _ = self._tcast(self.x16[j], rows, D)
                self._acc(f"mask_pred_list.{j}.k_proj.weight", self._wgrad(dkT, xT, D, D))
                if want_feat_grads:
                    d_f = torch.empty(rows, D, dtype=torch.float32, device=dev)
                    ops.linear(dk16, w["wkt"][j], d_f, M=rows, N=D, K=D)
                    d_feats[j] = d_f.view(B, S, D)
        for k, g in self.cls.grads.items():
            self.grads["cls_head." + k] = g
        return self.grads, d_feats


class _MaskHeadFunction(torch.autograd.Function):
    """Standalone training call of MaskHeadSegLevel (the one Query3DUnified makes after the decoder)."""

    @staticmethod
    def forward(ctx, mh, meta, query, *rest):
        n = len(mh.mask_pred_list)
        B, N, D = query.shape
        seed = None
        if mh.training and float(mh.cls_head[3].p) > 0.0:
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int32, device=query.device)
        t = MaskHeadTrain(mh, meta["feats"], meta["seg_masks"], B, N, seed)
        cls, logits, attn, sv = t.call(query.detach().reshape(B * N, D).float().contiguous())
        meta["attn"] = attn
        ctx.t, ctx.sv, ctx.mh, ctx.n = t, sv, mh, n
        ctx.in_dtype = query.dtype
        return cls, logits

    @staticmethod
    def backward(ctx, d_cls, d_logits):
        t, mh, n = ctx.t, ctx.mh, ctx.n
        B, N, D = t.B, t.N, t.D
        d_q = t.call_bwd(ctx.sv, d_cls, d_logits)
        grads, d_feats = t.finish(want_feat_grads=any(ctx.needs_input_grad[3:3 + n]))
        out = [None, None, None if d_q is None else d_q.view(B, N, D).to(ctx.in_dtype)]
        out += [d_feats[j] if ctx.needs_input_grad[3 + j] else None for j in range(n)]
        for name, p in mh.named_parameters():
            g = grads.get(name)
            out.append(None if g is None else g.reshape(p.shape).to(p.dtype))
        ctx.t = ctx.sv = None
        return tuple(out)
