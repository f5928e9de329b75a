"""Host-side mirror of the reference decoder modules (modules/grounding/query_encoder.py,
modules/layers/transformers.py:158-240) running on the sm_100a kernels behind the C ABI.

Same class names, constructor kwargs, parameter names/shapes (state_dict keys of SURVEY.md §8b) and
forward signatures as the reference, so a reference checkpoint loads with strict=True and
`cfg.model.unified_encoder.name: QueryMaskEncoder` resolves to this class when registered
(pq3d_b200/registry.py).  nn.Linear / nn.LayerNorm are used purely as parameter containers; all
arithmetic goes through `pq3d_b200.ops` (there is no PyTorch or CPU fallback).

Inference (eval mode, no autograd) replays one CUDA graph per forward.  With gradients enabled the forward and
backward run through pq3d_b200/train_engine.py (same kernels + backward.cu; dropout-free); what that path does
not cover raises instead of silently falling back.
"""
from __future__ import annotations

import math
import os
import weakref
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops

bf16 = torch.bfloat16


# --------------------------------------------------------------------------------------------
# parameter containers (identical names / shapes to the reference)
# --------------------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    """Parameters of nn.MultiheadAttention(d_model, nhead): packed in-projection + out_proj."""

    def __init__(self, d_model: int, nhead: int):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)


class CrossAttentionLayer(nn.Module):
    """modules/grounding/query_encoder.py:257-351 (post-norm; MHA with add_zero_attn=True)."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False, batch_first=False):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("pre-norm is never enabled on the reference path (query_encoder.py:97,103)")
        self.multihead_attn = _MHAParams(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class SelfAttentionLayer(nn.Module):
    """modules/grounding/query_encoder.py:184-254 (stock MHA self-attention, no zero-attn)."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False, batch_first=False):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("pre-norm is never enabled on the reference path")
        self.self_attn = _MHAParams(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class MultiHeadAttentionSpatial(nn.Module):
    """modules/layers/transformers.py:158-187, spatial_attn_fusion='mul', spatial_multihead=True,
    spatial_dim=5 — the only values reachable from SpatialSelfAttentionLayer's defaults."""

    def __init__(self, d_model, n_head, dropout=0.1, spatial_multihead=True, spatial_dim=5, spatial_attn_fusion="mul"):
        super().__init__()
        if spatial_attn_fusion != "mul" or not spatial_multihead or spatial_dim != 5:
            raise NotImplementedError("only fusion='mul', spatial_multihead=True, spatial_dim=5 are on the PQ3D path")
        self.n_head, self.d_model = n_head, d_model
        self.w_qs = nn.Linear(d_model, d_model)
        self.w_ks = nn.Linear(d_model, d_model)
        self.w_vs = nn.Linear(d_model, d_model)
        self.fc = nn.Linear(d_model, d_model)
        self.pairwise_loc_fc = nn.Linear(spatial_dim, n_head)


class SpatialSelfAttentionLayer(nn.Module):
    """modules/grounding/query_encoder.py:402-483."""

    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False, batch_first=False,
                 spatial_multihead=True, spatial_dim=5, spatial_attn_fusion="mul"):
        super().__init__()
        if normalize_before:
            raise NotImplementedError("pre-norm is never enabled on the reference path")
        self.self_attn = MultiHeadAttentionSpatial(d_model, nhead, dropout, spatial_multihead, spatial_dim,
                                                   spatial_attn_fusion)
        self.norm = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class FFNLayer(nn.Module):
    """modules/grounding/query_encoder.py:354-399 (relu, post-norm)."""

    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if activation != "relu" or normalize_before:
            raise NotImplementedError("the reference path uses relu / post-norm (query_encoder.py:97)")
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.dropout_p = dropout


class QueryEncoderLayer(nn.Module):
    """modules/grounding/query_encoder.py:96-112."""

    def __init__(self, d_model, nhead, memories, dim_feedforward=2048, dropout=0.1, activation="relu", prenorm=False,
                 spatial_selfattn=False, structure="mixed", memory_dropout=0, drop_memories_test=[]):
        super().__init__()
        sa = SpatialSelfAttentionLayer if spatial_selfattn else SelfAttentionLayer
        self.self_attn = sa(d_model, nhead, dropout=dropout, activation=activation, normalize_before=prenorm,
                            batch_first=True)
        self.cross_attn_list = nn.ModuleList(
            [CrossAttentionLayer(d_model, nhead, dropout=dropout, activation=activation, normalize_before=prenorm,
                                 batch_first=True) for _ in memories])
        self.ffn = FFNLayer(d_model, dim_feedforward, dropout=dropout, activation=activation, normalize_before=prenorm)
        self.structure = structure
        self.memories = list(memories)
        self.memory_dropout = memory_dropout
        self.drop_memories_test = list(drop_memories_test)
        if structure == "gate":
            self.gate_proj = nn.Linear(d_model, d_model)
        if structure not in ("sequential", "parallel", "mixed", "gate"):
            raise NotImplementedError(f"Unknow structure type: {structure}")


def _reference_init(encoder: "QueryMaskEncoder"):
    """Init parity: per-sublayer xavier on dim>1 params, deep-copied across
    memories and layers (modules/utils.py:28-32), then _init_weights_bert over every nn.Linear /
    nn.LayerNorm (modules/weights.py:3-19) — which leaves only the bare MHA in_proj_weight xavier."""
    template: Dict[Tuple[str, tuple], torch.Tensor] = {}
    with torch.no_grad():
        for name, p in encoder.named_parameters():
            if name.endswith("in_proj_weight"):
                kind = "self" if ".self_attn." in name else "cross"
                key = (kind, tuple(p.shape))
                if key not in template:
                    t = torch.empty_like(p)
                    nn.init.xavier_uniform_(t)
                    template[key] = t
                p.copy_(template[key])
            elif name.endswith("in_proj_bias"):
                p.zero_()
        for m in encoder.modules():
            if isinstance(m, nn.Linear):
                m.weight.normal_(mean=0.0, std=0.02)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.LayerNorm):
                m.bias.zero_()
                m.weight.fill_(1.0)


# --------------------------------------------------------------------------------------------
# packed weights
# --------------------------------------------------------------------------------------------
class _Packed:
    """bf16 operand copies of the fp32 parameters, laid out for the kernels:
       per memory  : Wk / Wv stacked over layers   [L*D, D]  (K/V projections hoisted out of the layer loop)
       per layer   : per CA group Wq stacked over the group's memories [g*D, D], out_proj [g*D, D],
                     LN gamma/beta [g, D]; self-attention [Wq;Wk] [2D, D], Wv, fc; FFN W1, W2.
    The buffers are allocated once; `refresh()` rewrites all of them from the live parameters with ONE kernel
    launch (pq3d_pack_segments reading a device-side segment table), so a training step re-packs for the cost of
    one pass over the weights, CUDA graphs that captured these pointers stay valid, and `train=True` adds the
    transposed copies the dgrad GEMMs take (`T(w)`)."""

    def __init__(self, enc: "QueryMaskEncoder", device, train: bool = False):
        D, L = enc.hidden_size, enc.num_layers
        layers = enc.unified_encoder
        mems = enc.memories
        self.device, self.train, self.stale = device, train, False
        self._order, self._fused_copy = list(mems), False
        self._segs: List[tuple] = []        # (src view, dst_c, dst_t, row0, fp32?)
        self._t: Dict[int, torch.Tensor] = {}
        self._keep: List[torch.Tensor] = []
        self._table = None

        w16, f = self._w16, self._f32
        self.wk, self.bk, self.wv, self.bv = {}, {}, {}, {}
        n_mem = len(mems)
        # every memory's Wk (Wv) in one buffer, in `memories` order: consecutive memories fuse into one grouped GEMM
        self._wk_all = self._alloc_w(n_mem * L * D, D) if n_mem else None
        self._wv_all = self._alloc_w(n_mem * L * D, D) if n_mem else None
        self._bk_all = torch.empty(n_mem * L * D, dtype=torch.float32, device=device)
        self._bv_all = torch.empty(n_mem * L * D, dtype=torch.float32, device=device)
        for j, m in enumerate(mems):
            ipw = [layers[i].cross_attn_list[j].multihead_attn.in_proj_weight for i in range(L)]
            ipb = [layers[i].cross_attn_list[j].multihead_attn.in_proj_bias for i in range(L)]
            sl = slice(j * L * D, (j + 1) * L * D)
            self.wk[m] = w16([w[D:2 * D] for w in ipw], into=self._wk_all, row0=j * L * D)
            self.wv[m] = w16([w[2 * D:] for w in ipw], into=self._wv_all, row0=j * L * D)
            self.bk[m] = f([b[D:2 * D] for b in ipb], into=self._bk_all[sl])
            self.bv[m] = f([b[2 * D:] for b in ipb], into=self._bv_all[sl])
        self.layers = []
        for i in range(L):
            lay = layers[i]
            d = {"groups": {}}
            for grp in enc._all_groups():
                idx = [mems.index(m) for m in grp]
                cas = [lay.cross_attn_list[j] for j in idx]
                d["groups"][grp] = dict(
                    wq=w16([c.multihead_attn.in_proj_weight[:D] for c in cas]),
                    bq=f([c.multihead_attn.in_proj_bias[:D] for c in cas]),
                    wo=w16([c.multihead_attn.out_proj.weight for c in cas]),
                    bo=f([c.multihead_attn.out_proj.bias for c in cas]).view(len(cas), D),
                    gamma=f([c.norm.weight for c in cas]).view(len(cas), D),
                    beta=f([c.norm.bias for c in cas]).view(len(cas), D),
                    eps=cas[0].norm.eps)
            sa = lay.self_attn
            if isinstance(sa, SpatialSelfAttentionLayer):
                a = sa.self_attn
                d["sa"] = dict(wqk=w16([a.w_qs.weight, a.w_ks.weight]), bqk=f([a.w_qs.bias, a.w_ks.bias]),
                               wv=w16([a.w_vs.weight]), bv=f([a.w_vs.bias]), wo=w16([a.fc.weight]), bo=f([a.fc.bias]),
                               loc_w=True, loc_b=True)
            else:
                a = sa.self_attn
                d["sa"] = dict(wqk=w16([a.in_proj_weight[:2 * D]]), bqk=f([a.in_proj_bias[:2 * D]]),
                               wv=w16([a.in_proj_weight[2 * D:]]), bv=f([a.in_proj_bias[2 * D:]]),
                               wo=w16([a.out_proj.weight]), bo=f([a.out_proj.bias]), loc_w=None, loc_b=None)
            d["sa"].update(gamma=f([sa.norm.weight])[None], beta=f([sa.norm.bias])[None], eps=sa.norm.eps)
            ffn = lay.ffn
            d["ffn"] = dict(w1=w16([ffn.linear1.weight]), b1=f([ffn.linear1.bias]), w2=w16([ffn.linear2.weight]),
                            b2=f([ffn.linear2.bias]), gamma=f([ffn.norm.weight])[None], beta=f([ffn.norm.bias])[None],
                            eps=ffn.norm.eps, F=ffn.linear1.out_features)
            if lay.structure == "gate":
                d["gate"] = dict(w=w16([lay.gate_proj.weight]), b=f([lay.gate_proj.bias]))
            self.layers.append(d)
        self._fused_kv = {}
        if self.layers[0]["sa"]["loc_w"] is not None:
            H = enc.num_heads
            self.loc_w = f([layers[i].self_attn.self_attn.pairwise_loc_fc.weight.view(-1) for i in range(L)]).view(L, H, 5)
            self.loc_b = f([layers[i].self_attn.self_attn.pairwise_loc_fc.bias for i in range(L)]).view(L, H)
            for i, d in enumerate(self.layers):
                d["sa"]["loc_w"], d["sa"]["loc_b"] = self.loc_w[i], self.loc_b[i]
        else:
            self.loc_w = self.loc_b = None
        self.refresh()

    # ---- registration ------------------------------------------------------------------------------
    def _alloc_w(self, rows, cols):
        w = torch.empty((rows, cols), dtype=bf16, device=self.device)
        if self.train:
            self._t[w.data_ptr()] = torch.empty((cols, rows), dtype=bf16, device=self.device)
            self._keep.append(w)
        return w

    def _src(self, p: torch.Tensor) -> torch.Tensor:
        p = p.detach()
        if p.dtype != torch.float32 or not ops.require_device_tensor(p) or not p.is_contiguous():
            raise TypeError("pq3d_b200: decoder parameters must be contiguous fp32 CUDA tensors (the kernels' bf16 "
                            f"operand copies are packed from them on the device); got {p.dtype} on {p.device}")
        return p

    def _w16(self, parts, into=None, row0=0):
        """bf16 copy of the row-wise concatenation of `parts` (2-D fp32 parameter views, same width)."""
        parts = [self._src(p) for p in parts]
        rows, cols = sum(p.shape[0] for p in parts), parts[0].shape[1]
        base = self._alloc_w(rows, cols) if into is None else into
        base_t = self._t.get(base.data_ptr())
        r = row0
        for p in parts:
            self._segs.append((p, base, base_t, r, False))
            r += p.shape[0]
        self._table = None
        return base if into is None else base[row0:row0 + rows]

    def _f32(self, parts, into=None):
        """fp32 copy of the concatenation of 1-D parameter views."""
        parts = [self._src(p) for p in parts]
        n = sum(p.numel() for p in parts)
        dst = torch.empty(n, dtype=torch.float32, device=self.device) if into is None else into
        o = 0
        for p in parts:
            self._segs.append((p.reshape(1, -1), dst[o:o + p.numel()].view(1, -1), None, 0, True))
            o += p.numel()
        self._table = None
        return dst

    def T(self, w: torch.Tensor, row0: int = 0, rows: Optional[int] = None) -> torch.Tensor:
        """Transposed bf16 copy [cols, rows] of a packed weight (train=True only); `row0/rows` select a row block of w,
        returned as the matching column block of the transpose."""
        t = self._t.get(w.data_ptr())
        if t is None:
            raise RuntimeError("transposed weight copies exist only for training (packed(train=True))")
        if t.shape[1] != w.shape[0]:            # w is a row slice of a larger buffer (wk[m] inside _wk_all)
            raise RuntimeError("T(): pass the base buffer and a row block")
        return t if rows is None else t[:, row0:row0 + rows]

    def T_mem(self, which: str, j: int, L: int, D: int) -> torch.Tensor:
        """Transpose [D, L*D] of memory j's stacked Wk / Wv."""
        base = self._wk_all if which == "k" else self._wv_all
        return self._t[base.data_ptr()][:, j * L * D:(j + 1) * L * D]

    # ---- the per-step refresh ------------------------------------------------------------------------
    def _build_table(self):
        rows, starts, total = [], [], 0
        for src, dst_c, dst_t, r0, is_f32 in self._segs:
            R, Cc = src.shape
            es = 4 if is_f32 else 2
            ld_c = dst_c.stride(0) if dst_c.ndim == 2 and dst_c.shape[0] > 1 else Cc
            pc = dst_c.data_ptr() + r0 * ld_c * es
            pt, ld_t = (0, 0) if dst_t is None else (dst_t.data_ptr() + r0 * 2, dst_t.stride(0))
            vec = (R % 4 == 0 and Cc % 4 == 0 and ld_c % 4 == 0 and ld_t % 4 == 0 and src.data_ptr() % 16 == 0
                   and pc % 16 == 0 and pt % 8 == 0)
            rows.append([src.data_ptr(), pc, pt, R, Cc, ld_c, ld_t, (1 if is_f32 else 0) | (2 if vec else 0)])
            starts.append(total)
            total += ((R + 63) // 64) * ((Cc + 63) // 64)
        self._table = (torch.tensor(rows, dtype=torch.int64).to(self.device),
                       torch.tensor(starts, dtype=torch.int32).to(self.device), total)

    def refresh(self):
        if self._table is None:
            self._build_table()
        ops.pack_segments(*self._table)
        self.stale = False

    def _fused(self, mems):
        """Stacked Wk / bk / Wv / bv of several memories (one grouped GEMM): a view when they are consecutive in the
        configured order, else an extra packed copy."""
        if mems not in self._fused_kv:
            order = self._order
            idx = [order.index(m) for m in mems]
            LD = self.wk[mems[0]].shape[0]
            if idx == list(range(idx[0], idx[0] + len(idx))):
                sl = slice(idx[0] * LD, (idx[-1] + 1) * LD)
                self._fused_kv[mems] = (self._wk_all[sl], self._bk_all[sl], self._wv_all[sl], self._bv_all[sl])
            else:
                self._fused_kv[mems] = (torch.cat([self.wk[m] for m in mems], 0), torch.cat([self.bk[m] for m in mems], 0),
                                        torch.cat([self.wv[m] for m in mems], 0), torch.cat([self.bv[m] for m in mems], 0))
                self._fused_copy = True
        return self._fused_kv[mems]

    fused_kv = _fused


class _MemState:
    """Per-forward projected state of one memory."""
    __slots__ = ("name", "S", "S_pitch", "multi", "per_layer", "xk", "xv", "K", "Vt", "bits", "strides", "tiles")


def _capturing() -> bool:
    """True while an outer CUDA-graph capture (Query3DUnified's whole-model graph) is recording this stream."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class _Branch:
    """A side stream for work that is independent of the critical chain (fork / join, capturable inside a CUDA graph):
    the small prologue kernels next to the memories' ingest + K / V^T projections, the self-attention V^T projection
    next to the q / k projection.  Disabled (everything in program order on the caller's stream) off-GPU."""

    def __init__(self, ws: dict, dev, enabled: bool):
        self.on = bool(enabled and dev.type == "cuda")
        self.dev = dev
        self.forked = False
        if self.on:
            self.stream = ws.get("side_stream")
            if self.stream is None:
                self.stream = ws["side_stream"] = torch.cuda.Stream(device=dev)

    def side(self):
        """Context: run on the side stream, after everything queued so far on the current stream (first use since the
        last join) — later uses keep queueing behind the side stream's own work."""
        import contextlib
        if not self.on:
            return contextlib.nullcontext()
        if not self.forked:
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
            self.forked = True
        return torch.cuda.stream(self.stream)

    def join(self):
        if self.on and self.forked:
            torch.cuda.current_stream(self.dev).wait_stream(self.stream)
            self.forked = False


# --------------------------------------------------------------------------------------------
# the decoder
# --------------------------------------------------------------------------------------------
class QueryMaskEncoder(nn.Module):
    """Drop-in for modules/grounding/query_encoder.py:51-94: same kwargs, same
    `.forward(input_dict, pairwise_locs, mask_head=None) -> (query, predictions_class, predictions_mask)`,
    `.spatial_selfattn` attribute read by the model (model/query3d_unified.py:182)."""

    def __init__(self, cfg=None, memories=[], memory_dropout=0.0, hidden_size=768, num_attention_heads=12,
                 num_layers=4, share_layer=False, spatial_selfattn=False, structure="sequential",
                 drop_memories_test=[], use_self_mask=False, num_blocks=1):
        super().__init__()
        if hidden_size % num_attention_heads != 0 or hidden_size // num_attention_heads != 64:
            raise NotImplementedError("the sm_100a attention kernel is specialised for head_dim = 64")
        if hidden_size % 128 != 0:
            raise NotImplementedError("hidden_size must be a multiple of 128")
        self.spatial_selfattn = spatial_selfattn
        memories = list(memories)
        if share_layer:
            layer = QueryEncoderLayer(hidden_size, num_attention_heads, memories, spatial_selfattn=spatial_selfattn,
                                      structure=structure, memory_dropout=memory_dropout,
                                      drop_memories_test=drop_memories_test)
            self.unified_encoder = nn.ModuleList([layer] * num_layers)
        else:
            self.unified_encoder = nn.ModuleList(
                [QueryEncoderLayer(hidden_size, num_attention_heads, memories, spatial_selfattn=spatial_selfattn,
                                   structure=structure, memory_dropout=memory_dropout,
                                   drop_memories_test=drop_memories_test) for _ in range(num_layers)])
        self.memories = memories
        self.structure = structure
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.memory_dropout = memory_dropout
        self.scene_meomories = [x for x in memories if x != "prompt"]     # (sic) reference attribute name
        self.drop_memories_test = list(drop_memories_test)
        self.use_self_mask = use_self_mask
        self.num_heads = num_attention_heads
        self.num_blocks = num_blocks
        _reference_init(self)
        self._packed: Optional[_Packed] = None
        self._packed_key = None
        self._packed_ver = None
        self._ws: Dict[tuple, dict] = {}
        self.use_cuda_graph = True
        # debugging / parity tap: a list here receives a clone of the fp32 query stream after every layer application
        # (forces the eager path; tests/test_parity_matched_gpu.py compares them layer by layer)
        self.layer_taps: Optional[List[torch.Tensor]] = None
        self.train_dropout = 0.1      # QueryEncoderLayer(dropout=0.1): sublayer / attention-probability / FFN dropout
        self._drop_seed = None
        self.last_memory_keep: List[torch.Tensor] = []
        # K / V^T of all layers are hoisted into one grouped GEMM per forward.  Projecting layer by layer (so a
        # layer's 75 MB stays L2-resident for its attention) was measured at config 3: the attention kernel did not
        # speed up (it is bound by TMEM->register bandwidth, not HBM) and the narrower GEMMs cost +63 us/step, so
        # the per-layer mode is off unless the hoisted buffers would exceed this many bytes.
        self.kv_hoist_bytes = 8 << 30
        # K / V^T of layer i+1 are query independent: project them on a SIDE stream (a fork inside the captured graph)
        # while layer i's latency-bound query-side chain (<= 48 CTAs per kernel) runs, with the GEMM's persistent grid
        # capped so the chain keeps SMs to run on.  One layer-sized K / V^T buffer (75 MB at config 3: L2-sized) is
        # reused by every layer.  Off: hoisted projection of all layers up front (num_blocks > 1 always hoists).
        # MEASURED (config 3 shard, B200, round 2): one batch at a time 0.773 -> 0.756 ms (-2 %; the background GEMM's
        # 100 CTAs and the chain's 48-144-CTA kernels slow each other down: out-projection 7 -> 24 us), but with four
        # batches in flight 0.503 -> 0.531 ms (+5 %: the narrower per-layer GEMMs are less efficient and the streams
        # already fill the machine).  Hence OFF by default; PQ3D_KV_OVERLAP=1 enables it for latency-bound serving.
        self.kv_overlap = os.environ.get("PQ3D_KV_OVERLAP", "0") != "0"
        self.kv_overlap_ctas = int(os.environ.get("PQ3D_KV_OVERLAP_CTAS", "100"))
        # with several batches in flight (one graph per stream) the persistent K / V^T projection of one batch would hold
        # every SM (232 KB of shared memory per CTA: nothing co-resides) while the other batches' latency-bound query-side
        # kernels wait; capping its grid leaves them room (0 = all SMs)
        self.shared_gemm_ctas = int(os.environ.get("PQ3D_SHARED_GEMM_CTAS", "0"))   # < 0: that many WAVES of CTAs
        # with batches in flight, every kernel but the long projections is launched at this CUDA priority (negative =
        # urgent; 0 = off): a stream's latency-bound chain then overtakes the other streams' projection CTAs
        self.inflight_priority = int(os.environ.get("PQ3D_INFLIGHT_PRIORITY", "0"))
        # with batches in flight the narrow-tile GEMMs keep 96 KB instead of 192 KB of operands in flight, so the
        # latency-bound query-side kernels of two streams share an SM (ops.shared_sm)
        self.inflight_half_ring = os.environ.get("PQ3D_INFLIGHT_HALF_RING", "1") != "0"
        # Opt-in: independent small kernels (mask packing, query copies, the prompt's projections, the spatial bias; the
        # self-attention V^T projection) on a side branch of the graph next to the critical chain.  MEASURED (round 2):
        # one batch at a time 0.7567 -> 0.7543 ms, no change with 4 batches in flight — and with 4 forked graphs in
        # flight the long (>= 10 k step) loops stalled on the device about once per 10-20 k replays (never with the
        # single-branch graphs: 30 k steps clean), so it stays OFF.
        self.side_branches = os.environ.get("PQ3D_SIDE_BRANCHES", "0") != "0"
        self.preingested = None        # set by Query3DUnified when its producers emitted the scene memories' operands
        # training backward: fork parameter-gradient work and the per-memory attention backwards onto side streams
        self.train_streams = True

    # ---- structure -> cross-attention program ------------------------------------------------
    def _active(self) -> List[str]:
        if self.training:
            return list(self.memories)                                              # training uses every memory, :156
        return [m for m in self.memories if m not in self.drop_memories_test]      # eval, :156

    def _group_is_parallel(self, gi: int) -> bool:
        """Whether group gi of `_program()` is a parallel_ca call (the only place memory dropout applies)."""
        return (self.structure == "parallel" or (self.structure == "mixed" and gi == 0)
                or (self.structure == "gate" and gi == 1))

    def _program(self) -> List[Tuple[str, ...]]:
        act = self._active()
        if self.structure == "sequential":
            return [(m,) for m in act]
        if self.structure == "parallel":
            assert "prompt" not in act
            return [tuple(act)]
        if self.structure == "mixed":
            return [tuple(m for m in act if m != "prompt"), ("prompt",)]
        if self.structure == "gate":
            return [("prompt",), tuple(m for m in self.memories if m != "prompt")]
        raise NotImplementedError(self.structure)

    def _all_groups(self):
        return list(dict.fromkeys(g for g in self._program() if len(g) > 0))

    # ---- packed weights ------------------------------------------------------------------------
    def packed(self, device, train: bool = False) -> _Packed:
        """The kernels' operand copies of the parameters: built once per (device, parameter storage), refreshed by
        one kernel launch whenever a parameter's version moved, after a training step (`stale`), or always in
        training (fused optimizers update parameters without bumping version counters)."""
        params = list(self.parameters())
        key = (str(device), tuple(self.drop_memories_test) if not self.training else (),
               tuple(p.data_ptr() for p in params))
        ver = tuple(p._version for p in params)
        pk = self._packed
        if pk is None or self._packed_key != key or (train and not pk.train):
            pk = self._packed = _Packed(self, device, train or (pk is not None and pk.train))
            self._packed_key, self._packed_ver = key, ver
            self._ws.clear()
        elif train or pk.stale or ver != self._packed_ver:
            if pk._fused_copy:           # torch.cat copies of non-consecutive fused memories: rebuild them
                pk._fused_kv.clear()
                self._ws.clear()
            pk.refresh()
            self._packed_ver = ver
        if train:
            pk.stale = True          # the optimizer step that follows leaves these copies one step behind
        return pk

    def mark_weights_changed(self):
        """Tell the decoder its parameters were updated behind autograd's back (CUDA-graph replay of an optimizer
        step): the next forward re-packs the operand copies."""
        if self._packed is not None:
            self._packed.stale = True

    def _buf(self, ws: dict, name: str, shape, dtype, device, zero: bool = False):
        t = ws.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            ws[name] = t
        return t

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, input_dict: dict, pairwise_locs: Optional[torch.Tensor], mask_head: Optional[Callable] = None):
        if self.training or (torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                                          or input_dict["query"][0].requires_grad)):
            # training path: forward (with the train-mode dropouts) + backward composed from the same kernels
            # (train_engine.py); scope limits raise.  module.train() under no_grad runs the same forward.
            from . import train_engine
            return train_engine.run(self, input_dict, pairwise_locs, mask_head)
        dev = input_dict["query"][0].device
        prio, inflight = 0, False
        if dev.type == "cuda":
            cur = torch.cuda.current_stream(dev).cuda_stream if getattr(self, "_stream_key_override", None) is None \
                else self._stream_key_override
            inflight = len({k[-2] for k in self._ws} | {cur}) > 1       # this decoder serves several streams
            if inflight:
                prio = self.inflight_priority
        with ops.launch_priority(prio), ops.shared_sm(inflight and self.inflight_half_ring):
            return self._forward_inference(input_dict, pairwise_locs, mask_head)

    def _forward_inference(self, input_dict: dict, pairwise_locs: Optional[torch.Tensor], mask_head: Optional[Callable]):
        query, query_masks, query_pos = input_dict["query"]
        dev = query.device
        B, N, D = query.shape
        H, L = self.num_heads, self.num_layers
        R = B * N
        pk = self.packed(dev)
        program = self._program()
        active = [m for g in program for m in g]
        # one workspace (static buffers + captured graph) per input-shape signature AND per CUDA stream, so several
        # batches can be in flight at once: a single batch's query-side chain is latency-bound and leaves most SMs
        # idle, two or three interleaved batches fill them (bench.py --streams)
        # (the captured body bakes in the mask buffers' pointers / strides and whether xk aliases xv, so mask rank and
        # the presence of a positional table are part of the key)
        pre = getattr(self, "preingested", None)
        key = (B, N, tuple((m, tuple(input_dict[m][0][0].shape if isinstance(input_dict[m][0], list)
                                     else input_dict[m][0].shape), input_dict[m][1].ndim, input_dict[m][2] is None)
                           for m in active),
               (torch.cuda.current_stream(dev).cuda_stream if getattr(self, "_stream_key_override", None) is None
                else self._stream_key_override) if dev.type == "cuda" else 0,
               None if pre is None else pre["xk"].data_ptr())
        ws = self._ws.setdefault(key, {})
        buf = lambda name, shape, dtype: self._buf(ws, name, shape, dtype, dev)  # noqa: E731

        if (self.use_cuda_graph and self.layer_taps is None and mask_head is None and not self.use_self_mask
                and ws.get("graph") is not None and not _capturing()):
            # Whole-forward graph: when the caller hands in the SAME device tensors again (a serving loop re-using its
            # staging buffers), the prologue that reads them (ingest, mask packing, copies) is captured together with
            # the body, so a forward costs one graph launch on the host.  Fresh tensors fall back to the eager prologue
            # + body graph below; a signature is captured the second time it is seen, at most 8 are kept.
            sig, tensors = self._input_signature(input_dict, pairwise_locs)
            full = ws.setdefault("full", {})
            ent = full.get(sig)
            if ent is not None and ent.get("graph") is None and not all(r() is t for r, t in zip(ent["refs"], tensors)):
                ent = None                 # same addresses, different tensors (allocator re-use): not a stable buffer
            if ent is not None and ent.get("graph") is not None:
                ent["graph"].replay()
                ops._count(ent["launches"])
                if "voxel" in input_dict and isinstance(input_dict["voxel"][0], list):
                    input_dict["voxel"][0] = input_dict["voxel"][0][L - 1]
                return ent["out"].clone(), [], []
            if ent is not None:
                torch.cuda.synchronize()
                before = ops.LAUNCHES
                g = torch.cuda.CUDAGraph()
                ws["_eager_body"] = True
                try:
                    with torch.cuda.graph(g):
                        out = self._forward_core(dict(input_dict), pairwise_locs, None, ws, pk, program, active)[0]
                finally:
                    ws["_eager_body"] = False
                ent.update(graph=g, out=out, launches=ops.LAUNCHES - before, keep=tensors)
                ops.LAUNCHES = before
                g.replay()
                ops._count(ent["launches"])
                if "voxel" in input_dict and isinstance(input_dict["voxel"][0], list):
                    input_dict["voxel"][0] = input_dict["voxel"][0][L - 1]
                return out.clone(), [], []
            if len(full) >= 8:
                full.pop(next(iter(full)))
            full[sig] = {"refs": [weakref.ref(t) for t in tensors]}
        return self._forward_core(input_dict, pairwise_locs, mask_head, ws, pk, program, active)

    @staticmethod
    def _flat_tensors(v):
        if isinstance(v, torch.Tensor):
            return [v]
        if isinstance(v, (list, tuple)):
            return [t for x in v for t in QueryMaskEncoder._flat_tensors(x)]
        return []

    def _input_signature(self, input_dict, pairwise_locs):
        """(hashable signature, tensors): address, shape, strides and dtype of every tensor the forward reads."""
        tensors = [] if pairwise_locs is None else [pairwise_locs]
        for k in sorted(input_dict):
            tensors += self._flat_tensors(input_dict[k])
        return tuple((t.data_ptr(), tuple(t.shape), t.stride(), t.dtype) for t in tensors), tensors

    def _forward_core(self, input_dict, pairwise_locs, mask_head, ws, pk, program, active):
        query, query_masks, query_pos = input_dict["query"]
        dev = query.device
        B, N, D = query.shape
        H, L = self.num_heads, self.num_layers
        R = B * N
        buf = lambda name, shape, dtype: self._buf(ws, name, shape, dtype, dev)  # noqa: E731

        # ---------------- prologue (eager): everything that reads caller-owned tensors lands in static buffers
        voxel_feat = input_dict["voxel"][0] if "voxel" in input_dict else None
        states: Dict[str, _MemState] = {}
        # memories with one feature tensor, a positional table and the same token count share one set of
        # buffers so that their K / V^T projections run as ONE grouped GEMM launch each
        fused = [m for m in active if not isinstance(input_dict[m][0], list) and input_dict[m][2] is not None]
        if len(fused) >= 2 and len({input_dict[m][0].shape[1] for m in fused}) == 1:
            nf = len(fused)
            Sp = ops.pad8(input_dict[fused[0]][0].shape[1])
            tag = "+".join(fused)
            # the model's producers may already have written these memories' bf16 operands (Query3DUnified, §8f-1)
            pre = getattr(self, "preingested", None)
            self.preingested = None
            if not (pre is not None and pre["mems"] == tuple(fused) and pre["S"] == Sp and pre["B"] == B
                    and all(input_dict[m][2] is pre["pos"] for m in fused)):
                pre = None
            if pre is not None:
                xk_all, xv_all = pre["xk"], pre["xv"]
            else:
                xk_all, xv_all = buf(f"xk_{tag}", (nf * B * Sp, D), bf16), buf(f"xv_{tag}", (nf * B * Sp, D), bf16)
            # K / V^T of all L layers for these memories: nf*B*Sp*L*D*4 bytes (302 MB at config 3); past
            # kv_hoist_bytes project layer by layer into one layer-sized buffer instead (bounds memory).
            overlap = bool(self.kv_overlap and self.num_blocks == 1 and L > 1 and dev.type == "cuda")
            per_layer = self.num_blocks == 1 and (overlap or nf * B * Sp * L * D * 4 > self.kv_hoist_bytes)
            Lk = 1 if per_layer else L
            K_all, Vt_all = buf(f"K_{tag}", (nf, B * Sp, Lk * D), bf16), buf(f"Vt_{tag}", (nf, Lk * D, B * Sp), bf16)
        else:
            fused, per_layer, overlap, pre = [], False, False, None
            self.preingested = None
        br = _Branch(ws, dev, self.side_branches)
        ws["_branch"] = br
        import contextlib
        # the fused memories usually share ONE positional table object (fts_pos): ingest them in one launch
        fused_shared_pos = bool(fused) and len(fused) <= 4 and all(input_dict[m][2] is input_dict[fused[0]][2] for m in fused)
        for m in active:
            feat, mask, pos = input_dict[m]
            multi = isinstance(feat, list)
            f0 = feat[0] if multi else feat
            S = f0.shape[1]
            Sp = ops.pad8(S)
            st = _MemState()
            st.name, st.S, st.S_pitch, st.multi = m, S, Sp, multi
            st.per_layer = per_layer and m in fused
            nsrc = L if multi else 1
            if m in fused:
                j = fused.index(m)
                st.xv, st.xk = xv_all[j * B * Sp:(j + 1) * B * Sp], xk_all[j * B * Sp:(j + 1) * B * Sp]
                st.K, st.Vt = K_all[j], Vt_all[j]
            else:
                st.xv = buf(f"xv_{m}", (nsrc * B * Sp, D), bf16)
                st.xk = buf(f"xk_{m}", (nsrc * B * Sp, D), bf16) if pos is not None else st.xv
                st.K = buf(f"K_{m}", (B * Sp, L * D), bf16)
                st.Vt = buf(f"Vt_{m}", (L * D, B * Sp), bf16)
            # the big memories' ingest (and below their K / V^T projections) form the critical chain; everything else of
            # the prologue is independent of it and goes to the side branch
            shared = m in fused and (fused_shared_pos or pre is not None)
            if shared and m == fused[0] and pre is None:  # one launch for all fused memories: pos is read once
                fl = [input_dict[x][0].contiguous() for x in fused]
                ops.ingest_memories([t if t.dtype == torch.float32 else t.float() for t in fl], pos.contiguous().float(),
                                    xk_all, xv_all, B * Sp * D, Sp)
            with (contextlib.nullcontext() if (m in fused or not fused) else br.side()):
                for i in range(0 if shared else nsrc):
                    fi = (feat[i] if multi else feat).contiguous()
                    sl = slice(i * B * Sp, (i + 1) * B * Sp)
                    ops.ingest_memory(fi.float() if fi.dtype != torch.float32 else fi,
                                      None if pos is None else pos.contiguous().float(),
                                      st.xk[sl] if pos is not None else None, st.xv[sl], Sp)
            with br.side():
                self._set_mask(st, mask, B, N, H, ws, dev)
            states[m] = st
        q32 = buf("q32", (R, D), torch.float32)
        qpos = buf("qpos", (R, D), torch.float32)
        xq = buf("xq", (R, D), bf16)       # bf16(query + query_pos): q/k operand
        xv_q = buf("xvq", (R, D), bf16)    # bf16(query): v operand
        if self.spatial_selfattn and pairwise_locs is None:
            raise ValueError("spatial_selfattn=True needs pairwise_locs (B, N, N, 5)")
        pw = buf("pw", (B, N, N, 5), torch.float32) if self.spatial_selfattn else None
        with br.side():
            q32.copy_(query.reshape(R, D))
            qpos.copy_(query_pos.reshape(R, D))
            qbits = ops.pack_mask(query_masks.contiguous(), buf("qbits", (B, ops.mask_words(N)), torch.int32))
            if pw is not None:
                pw.copy_(pairwise_locs)
        br.join()          # (the eager prologue and the captured body are separate stream segments)
        # spatial score bias of all L layers, computed once per forward from the (layer independent) geometry
        sbias = buf("sbias", (L, B, H, N, ops.bias_ld(N)), torch.float32) if pk.loc_w is not None else None

        def project_fused(l0, nl, max_ctas=0):
            """K / V^T of layers [l0, l0+nl) for the fused memories: one grouped GEMM launch each."""
            nf, Sp = len(fused), states[fused[0]].S_pitch
            wk, bk, wv, bv = pk.fused_kv(tuple(fused))
            # CTA-pair GEMMs only while this decoder works on ONE stream: with pair GEMMs of several in-flight graphs
            # competing for SMs, long loops stalled on the device (csrc/gemm.cu, flags bit 1)
            solo = len({k[-2] for k in self._ws}) <= 1
            if not solo and max_ctas == 0:
                max_ctas = self.shared_gemm_ctas      # batches in flight: leave SMs to the other streams' small kernels
            with ops.launch_priority(0):              # the long GEMMs yield to every urgent kernel (inflight_priority)
                ops.linear(xk_all, wk[l0 * D:], K_all, M=B * Sp, N=nl * D, K=D, bias=bk[l0 * D:], bias_group_stride=L * D,
                           groups=nf, a_group_rows=B * Sp, w_group_rows=L * D, ldc=nl * D, c_group_stride=B * Sp * nl * D,
                           max_ctas=max_ctas, no_pairs=not solo)
                ops.linear(wv[l0 * D:], xv_all, Vt_all, M=nl * D, N=B * Sp, K=D, bias=bv[l0 * D:], bias_along_m=True,
                           bias_group_stride=L * D, groups=nf, a_group_rows=L * D, w_group_rows=B * Sp, ldc=B * Sp,
                           c_group_stride=nl * D * B * Sp, max_ctas=max_ctas, no_pairs=not solo)

        def project_memories():
            """Hoisted K / V^T projections of every memory for all L layers (query independent)."""
            # fork first: the small projections / casts / spatial bias must not queue behind the big GEMMs
            with (br.side() if fused else contextlib.nullcontext()):
                project_small()
            if fused and not per_layer:
                project_fused(0, L)
            br.join()

        def project_small():
            for m in active:
                if m in fused:
                    continue
                st = states[m]
                Sp = st.S_pitch
                if st.multi:
                    ops.linear(st.xk, pk.wk[m], st.K, M=B * Sp, N=D, K=D, bias=pk.bk[m], bias_group_stride=D,
                               groups=L, a_group_rows=B * Sp, w_group_rows=D, ldc=L * D, c_group_stride=D)
                    ops.linear(pk.wv[m], st.xv, st.Vt, M=D, N=B * Sp, K=D, bias=pk.bv[m], bias_along_m=True,
                               bias_group_stride=D, groups=L, a_group_rows=D, w_group_rows=B * Sp, ldc=B * Sp,
                               c_group_stride=D * B * Sp)
                else:
                    solo = len({k[-2] for k in self._ws}) <= 1
                    ops.linear(st.xk, pk.wk[m], st.K, M=B * Sp, N=L * D, K=D, bias=pk.bk[m], no_pairs=not solo)
                    ops.linear(pk.wv[m], st.xv, st.Vt, M=L * D, N=B * Sp, K=D, bias=pk.bv[m], bias_along_m=True,
                               no_pairs=not solo)
            ops.cast_bf16(q32, xq, add=qpos)
            ops.cast_bf16(q32, xv_q)
            if sbias is not None:
                ops.spatial_bias(pw, pk.loc_w, pk.loc_b, sbias)

        side = None
        if overlap:
            side = ws.get("kv_side")
            if side is None:
                side = ws["kv_side"] = torch.cuda.Stream(device=dev)

        def kv_ready(i):
            """Called right before layer i's first attention over the fused memories."""
            if not per_layer:
                return
            if not overlap:
                project_fused(i, 1)
            elif i == 0:
                project_fused(0, 1)
            else:
                torch.cuda.current_stream(dev).wait_stream(side)          # join: K / V^T of layer i are complete

        def kv_done(i):
            """Called right after layer i's last attention over the fused memories: their buffer is free again."""
            if overlap and i + 1 < L:
                side.wait_stream(torch.cuda.current_stream(dev))          # fork
                with torch.cuda.stream(side):
                    project_fused(i + 1, 1, max_ctas=self.kv_overlap_ctas)

        def run_layer(i):
            self._layer(i, pk, program, states, ws, dev, B, N, D, H, q32, qpos, xq, xv_q, qbits, sbias,
                        fused=fused, kv_ready=kv_ready, kv_done=kv_done)
            if self.layer_taps is not None:
                self.layer_taps.append(q32.view(B, N, D).clone())

        predictions_class, predictions_mask = [], []
        if mask_head is None and not self.use_self_mask:
            # ---------------- body touches static buffers only: replay it as one CUDA graph
            def body():
                project_memories()
                for _block in range(self.num_blocks):
                    for i in range(L):
                        run_layer(i)
            self._run_body(ws, body)
            if isinstance(voxel_feat, list):
                input_dict["voxel"][0] = voxel_feat[L - 1]
            return q32.view(B, N, D).clone(), predictions_class, predictions_mask

        mh = self._own_mask_head(mask_head)
        if mh is not None:
            # ---------------- our own MaskHeadSegLevel in the loop: its hoisted part runs once here, its per-layer
            # part writes static buffers, so blocks x layers x (mask head + mask packing + layer) is one CUDA graph
            kw = mask_head.keywords
            slot = ("decoder", id(ws))
            mh.prepare(kw["seg_fts_for_match"], kw["seg_masks"], slot)
            S_m, C = kw["seg_masks"].shape[1], mh.cls_head[4].out_features
            n_calls = self.num_blocks * L
            cls_buf = buf("mh_cls", (n_calls, B, N, C), torch.float32)
            logit_buf = buf("mh_logits", (n_calls, B, S_m, N), torch.float32)
            attn_buf = buf("mh_attn", (B, N, S_m), torch.bool)
            fixed = buf("am_fixed", (B, N, S_m), torch.bool)
            am_bits = buf("am_bits", (B, N, ops.mask_words(S_m)), torch.int32)
            am_tiles = buf("am_tiles", (B,), torch.int32)

            def body():
                project_memories()
                for k in range(n_calls):
                    mh.run_into(q32, B, N, cls_buf[k], logit_buf[k], attn_buf, slot)
                    if self.use_self_mask:
                        ops.pack_mask(attn_buf, am_bits, unmask_full_rows=True, mask_fixed=fixed.view(torch.uint8),
                                      active_tiles=am_tiles)
                        for st in states.values():
                            if st.name != "prompt":
                                st.bits, st.strides, st.tiles = am_bits, (am_bits.stride(0), 0, am_bits.stride(1)), am_tiles
                    run_layer(k % L)
            self._run_body(ws, body)
            predictions_class = [cls_buf[k].clone() for k in range(n_calls)]
            predictions_mask = [logit_buf[k].clone() for k in range(n_calls)]
            if self.use_self_mask:
                # what the reference leaves behind (query_encoder.py:84-88): attn_mask.repeat_interleave(H, 0),
                # shape (B*H, N, S), row b*H + h — materialised once per forward, outside the graph
                out_mask = fixed.repeat_interleave(H, 0)
                for m in input_dict.keys():
                    if m not in ("query", "prompt"):
                        input_dict[m][1] = out_mask
            if isinstance(voxel_feat, list):
                input_dict["voxel"][0] = voxel_feat[L - 1]
            return q32.view(B, N, D).clone(), predictions_class, predictions_mask

        # ---------------- generic mask_head callable: eager loop
        project_memories()
        attn_mask = None
        for _block in range(self.num_blocks):
            for i in range(L):
                if mask_head is not None:
                    output_class, outputs_mask, attn_mask = mask_head(q32.view(B, N, D))
                    predictions_class.append(output_class)
                    predictions_mask.append(outputs_mask)
                if self.use_self_mask:
                    if attn_mask is None:
                        raise ValueError("use_self_mask=True needs a mask_head that returns an attention mask")
                    fixed = buf("am_fixed", (B, N, attn_mask.shape[-1]), torch.bool)
                    am_tiles = buf("am_tiles", (B,), torch.int32)
                    bits = ops.pack_mask(attn_mask.contiguous(),
                                         buf("am_bits", (B, N, ops.mask_words(attn_mask.shape[-1])), torch.int32),
                                         unmask_full_rows=True, mask_fixed=fixed.view(torch.uint8),
                                         active_tiles=am_tiles)
                    rep = None
                    for m in input_dict.keys():
                        if m in ("query", "prompt"):
                            continue
                        # the reference stores attn_mask.repeat_interleave(H, 0) here (query_encoder.py:84-88); the
                        # kernels read the packed (B, N, S) bits with a zero head stride instead
                        if rep is None:
                            rep = fixed.repeat_interleave(H, 0)
                        input_dict[m][1] = rep
                        if m in states:
                            st = states[m]
                            st.bits, st.strides, st.tiles = bits, (bits.stride(0), 0, bits.stride(1)), am_tiles
                if isinstance(voxel_feat, list):
                    input_dict["voxel"][0] = voxel_feat[i]
                run_layer(i)
        return q32.view(B, N, D).clone(), predictions_class, predictions_mask

    @staticmethod
    def _own_mask_head(mask_head):
        """Our MaskHeadSegLevel wired the way Query3DUnified wires it (functools.partial with keyword tensors,
        predictions enabled, no offline masks) -> the module, else None."""
        import functools
        from .mask_head import MaskHeadSegLevel
        if not isinstance(mask_head, functools.partial) or mask_head.args:
            return None
        owner = getattr(mask_head.func, "__self__", mask_head.func)
        kw = mask_head.keywords
        if (not isinstance(owner, MaskHeadSegLevel) or kw.get("skip_prediction", False)
                or kw.get("offline_attn_masks") is not None or "seg_fts_for_match" not in kw or "seg_masks" not in kw):
            return None
        return owner

    def _run_body(self, ws: dict, body: Callable):
        """First call per shape: eager (allocates the workspace, configures kernels).  Second call:
        capture.  Afterwards: one graph launch per forward (the launch-bound query-side chain of ~70
        small kernels costs more on the host than on the GPU otherwise)."""
        if (not self.use_cuda_graph or ws.get("_eager_body") or self.layer_taps is not None
                or _capturing()):                                 # an outer capture (whole-model graph) records the launches
            body()
            return
        g = ws.get("graph")
        if g is not None:
            g.replay()
            ops._count(ws["graph_launches"])
            return
        if not ws.get("warm"):
            body()
            ws["warm"] = True
            return
        torch.cuda.synchronize()
        before = ops.LAUNCHES
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        ws["graph_launches"] = ops.LAUNCHES - before
        ops.LAUNCHES = before
        ws["graph"] = g
        g.replay()
        ops._count(ws["graph_launches"])

    def _set_mask(self, st: _MemState, mask: torch.Tensor, B, N, H, ws, dev):
        if mask.dtype != torch.bool:
            raise TypeError(f"memory '{st.name}': masks must be torch.bool (True = ignore), got {mask.dtype}")
        W = ops.mask_words(st.S)
        if mask.ndim == 2:                                    # key padding (B, S)
            if tuple(mask.shape) != (B, st.S):
                raise ValueError(f"memory '{st.name}': key padding mask {tuple(mask.shape)} != {(B, st.S)}")
            st.tiles = self._buf(ws, f"tiles2_{st.name}", (B,), torch.int32, dev)
            st.bits = ops.pack_mask(mask.contiguous(), self._buf(ws, f"bits2_{st.name}", (B, W), torch.int32, dev),
                                    active_tiles=st.tiles)
            st.strides = (W, 0, 0)
        elif mask.ndim == 3:                                  # attn mask (B*H, N, S), row b*H + h
            if tuple(mask.shape) != (B * H, N, st.S):
                raise RuntimeError(f"The shape of the 3D attn_mask is {tuple(mask.shape)}, but should be {(B * H, N, st.S)}.")
            st.tiles = None      # per-head masks: (B*H) leading entries; tail trimming not wired for this layout
            st.bits = ops.pack_mask(mask.contiguous(), self._buf(ws, f"bits3_{st.name}", (B * H, N, W), torch.int32, dev))
            st.strides = (H * N * W, N * W, W)
        else:
            raise ValueError(f"memory '{st.name}': mask must be 2-D or 3-D")

    def _cross_group(self, grp, i, pk, states, ws, dev, B, N, D, H, x_in, tag, after_attention=None):
        """Q projection for the group's memories, one attention launch, grouped out-projection.
        Returns y [g, R, D] fp32 (pre-residual updates)."""
        R, g = B * N, len(grp)
        w = pk.layers[i]["groups"][grp]
        Q = self._buf(ws, f"Q_{tag}", (R, g * D), bf16, dev)
        ops.linear(x_in, w["wq"], Q, M=R, N=g * D, K=D, bias=w["bq"], alpha=ops.Q_SCALE, alpha_ncols=g * D, w_const=True)
        O = self._buf(ws, f"O_{tag}", (g, R, D), bf16, dev)
        mems = [ops.AttnMemory(states[m].K, 0 if states[m].per_layer else i * D, states[m].Vt,
                               0 if states[m].per_layer else i * D, states[m].S, states[m].S_pitch, states[m].bits,
                               *states[m].strides, kv_tiles=states[m].tiles) for m in grp]
        ops.attention(Q, D, mems, O, R * D, B, H, N, True)
        if after_attention is not None:
            after_attention()
        y = self._buf(ws, f"y_{tag}", (g, R, D), torch.float32, dev)
        ops.linear(O.view(g * R, D), w["wo"], y, M=R, N=D, K=D, bias=w["bo"], bias_group_stride=D, groups=g,
                   a_group_rows=R, w_group_rows=D, ldc=D, c_group_stride=R * D, w_const=True)
        return y, w

    def _layer(self, i, pk, program, states, ws, dev, B, N, D, H, q32, qpos, xq, xv_q, qbits, sbias, fused=(),
               kv_ready=None, kv_done=None):
        R = B * N
        lw = pk.layers[i]
        # groups that attend the fused (per-layer projected) memories: K / V^T must be ready before the first of them
        # and their buffer is released after the last one
        uses = [gi for gi, grp in enumerate(program) if any(m in fused for m in grp)]
        first_use, last_use = (uses[0], uses[-1]) if uses else (-1, -1)

        def cross(gi, grp, x_in, tag):
            if gi == first_use and kv_ready is not None:
                kv_ready(i)
            out = self._cross_group(grp, i, pk, states, ws, dev, B, N, D, H, x_in, tag,
                                    after_attention=(lambda: kv_done(i)) if (gi == last_use and kv_done is not None) else None)
            return out
        if self.structure == "gate":
            grp_p, grp_s = program
            y, w = cross(0, grp_p, xq, "p")
            pb = self._buf(ws, "gate_p16", (R, D), bf16, dev)
            ops.add_layernorm(y, q32, w["gamma"], w["beta"], w["eps"], R, D, G=1, out_bf16=pb)
            gl = self._buf(ws, "gate_logits", (R, D), torch.float32, dev)
            ops.linear(pb, lw["gate"]["w"], gl, M=R, N=D, K=D, bias=lw["gate"]["b"], w_const=True)
            y, w = cross(1, grp_s, xq, "s")
            upd = self._buf(ws, "gate_upd", (R, D), torch.float32, dev)
            ops.add_layernorm(y, q32, w["gamma"], w["beta"], w["eps"], R, D, G=len(grp_s), y_group_stride=R * D,
                              out_f32=upd)
            ops.gate_mix(gl, q32, upd, q32)
            ops.cast_bf16(q32, xq, add=qpos)
            ops.cast_bf16(q32, xv_q)
        else:
            for gi, grp in enumerate(program):
                if len(grp) == 0:
                    continue
                y, w = cross(gi, grp, xq, f"g{gi}")
                ops.add_layernorm(y, q32, w["gamma"], w["beta"], w["eps"], R, D, G=len(grp), y_group_stride=R * D,
                                  pos=qpos, out_f32=q32, out_bf16=xv_q, out_pos_bf16=xq)
        # ---- query self-attention (spatially biased when configured)
        sa = lw["sa"]
        QK = self._buf(ws, "sa_QK", (R, 2 * D), bf16, dev)
        ops.linear(xq, sa["wqk"], QK, M=R, N=2 * D, K=D, bias=sa["bqk"], alpha=ops.Q_SCALE, alpha_ncols=D, w_const=True)
        # V^T [D, B*Np]: scene b's queries at columns b*Np .. b*Np+N (Np = N rounded up to 8 for the TMA
        # stride); one GEMM group per scene, pad columns stay at their zero initialisation
        Np = ops.pad8(N)
        Vt = self._buf(ws, "sa_Vt", (D, B * Np), bf16, dev, zero=True)
        br = ws.get("_branch")
        import contextlib
        with (br.side() if br is not None else contextlib.nullcontext()):     # independent of the q / k projection above
            ops.linear(sa["wv"], xv_q, Vt, M=D, N=N, K=D, bias=sa["bv"], bias_along_m=True, groups=B, a_group_rows=0,
                       w_group_rows=N, ldc=B * Np, c_group_stride=Np)
        if br is not None:
            br.join()
        Os = self._buf(ws, "sa_O", (1, R, D), bf16, dev)
        mem = ops.AttnMemory(QK, D, Vt, 0, N, N, qbits, qbits.stride(0), 0, 0, Vt_pitch=Np)
        ops.attention(QK, 0, [mem], Os, R * D, B, H, N, False, None if sbias is None else sbias[i])
        ys = self._buf(ws, "sa_y", (R, D), torch.float32, dev)
        ops.linear(Os.view(R, D), sa["wo"], ys, M=R, N=D, K=D, bias=sa["bo"], w_const=True)
        ops.add_layernorm(ys, q32, sa["gamma"], sa["beta"], sa["eps"], R, D, out_f32=q32, out_bf16=xv_q)
        # ---- FFN
        ff = lw["ffn"]
        h = self._buf(ws, "ffn_h", (R, ff["F"]), bf16, dev)
        ops.linear(xv_q, ff["w1"], h, M=R, N=ff["F"], K=D, bias=ff["b1"], relu=True, w_const=True)
        yf = self._buf(ws, "ffn_y", (R, D), torch.float32, dev)
        ops.linear(h, ff["w2"], yf, M=R, N=D, K=ff["F"], bias=ff["b2"], w_const=True)
        ops.add_layernorm(yf, q32, ff["gamma"], ff["beta"], ff["eps"], R, D, pos=qpos, out_f32=q32, out_bf16=xv_q,
                          out_pos_bf16=xq)


class QueryEncoder(QueryMaskEncoder):
    """modules/grounding/query_encoder.py:11-49: the mask-less variant (`forward(input_dict, pairwise_locs) -> query`).
    Its layers are built WITHOUT memory_dropout / drop_memories_test (:18); the class drops memories itself
    (`dropout_memory`, :26-36): in training each scene's feat / pos of every scene memory is zeroed with probability
    `memory_dropout`, in eval the memories in `drop_memories_test` are zeroed for every scene — the memory is still
    attended (zero keys / values), unlike QueryMaskEncoder's layer-level drop."""

    def __init__(self, cfg=None, memories=[], memory_dropout=0.0, hidden_size=768, num_attention_heads=12,
                 num_layers=4, share_layer=False, spatial_selfattn=False, structure="sequential",
                 drop_memories_test=[]):
        super().__init__(cfg, memories, 0.0, hidden_size, num_attention_heads, num_layers, share_layer,
                         spatial_selfattn, structure, [], False, 1)
        self.memory_dropout = memory_dropout          # the reference attribute (:22); layers keep memory_dropout = 0
        self.layer_memory_dropout = 0.0
        self.drop_memories_test = list(drop_memories_test)
        self.last_scene_drop: Dict[str, torch.Tensor] = {}

    def _active(self) -> List[str]:
        return list(self.memories)                    # layer-level drop_memories_test is empty for this class (:18)

    def dropout_memory(self, input_dict):
        """:26-36, in place on the caller's tensors like the reference."""
        self.last_scene_drop = {}
        for memory in self.scene_meomories:
            if memory not in input_dict:
                continue
            feat, mask, pos = input_dict[memory]
            nb = feat.shape[0]
            if self.training:
                drop_mask = torch.rand(nb, device=feat.device) < self.memory_dropout
            elif memory in self.drop_memories_test:
                drop_mask = torch.ones(nb, device=feat.device, dtype=torch.bool)
            else:
                drop_mask = torch.zeros(nb, device=feat.device, dtype=torch.bool)
            self.last_scene_drop[memory] = drop_mask
            feat[drop_mask] = 0.
            if pos is not None:
                pos[drop_mask] = 0.

    def forward(self, input_dict, pairwise_locs):
        if (self.training and self.memory_dropout > 0) or (not self.training and self.drop_memories_test):
            self.dropout_memory(input_dict)
        return super().forward(input_dict, pairwise_locs, None)[0]
