"""Training path of the decoder: forward that keeps what the backward needs, and the backward itself,
both composed from the sm_100a kernels behind the C ABI (no torch autograd through PyTorch ops, no fallback).

What the reference does here is plain autograd through modules/grounding/query_encoder.py (trainer/
query3d_trainer.py:18-28: forward, loss, backward, AdamW).  This module gives `QueryMaskEncoder` the same
contract — `enc(input_dict, pairwise_locs)` under grad mode returns a tensor attached to the graph, and
`loss.backward()` fills `.grad` of every decoder parameter plus the gradients of the query / memory inputs —
through ONE `torch.autograd.Function` whose backward runs:

  GEMM-shaped work   dgrad  d_x = d_y W        -> pq3d_linear_bf16 with a transposed bf16 weight copy
                     wgrad  dW = d_y^T x        -> pq3d_linear_bf16 on operands transposed by pq3d_transpose_cast
                     attention: S2 recompute, dP = dO V^T, dV = P^T dO, dK = dS^T Q, dQ = dS K -> pq3d_bgemm_bf16
  streaming work     LayerNorm backward, softmax backward (both orientations), bias column sums, the
                     spatial-bias backward -> backward.cu

Scope: structures sequential / parallel / mixed / gate, any num_blocks, multi-scale (per-layer) memory
features, the in-loop mask head (our MaskHeadSegLevel) with use_self_mask, train-mode dropout and memory dropout.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple

import torch

from . import ops, rng

bf16 = torch.bfloat16
f32 = torch.float32


def _e(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def _z(shape, dtype, dev):
    return torch.zeros(shape, dtype=dtype, device=dev)


class _Mem:
    __slots__ = ("name", "S", "Sp", "xk", "xv", "K", "Vt", "bits", "strides", "tiles", "has_pos", "S_pitch", "multi")


def _heads(t2d: torch.Tensor, B: int, rows: int, pitch: int, H: int, col0: int = 0) -> torch.Tensor:
    """(B*pitch, ld) row-major tensor -> strided (B, H, rows, 64) view of columns [col0, col0 + H*64)."""
    ld = t2d.stride(0)
    return t2d.as_strided((B, H, rows, 64), (pitch * ld, 64, ld, 1), t2d.storage_offset() + col0)


_SIDE: Dict = {}


def dist_ready() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _side_streams(dev, n):
    """Per-device pool of side streams: work off the backward's critical path (weight / bias gradients, operand
    transposes) and the independent per-memory attention backwards run next to the main chain."""
    key = (dev.type, dev.index)
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def forward(enc, query, query_pos, query_masks, mems: List[Tuple[str, torch.Tensor, torch.Tensor, Optional[torch.Tensor]]],
            pairwise_locs, mh=None, mh_kw=None):
    """mems: [(name, feat (B,S,D), mask bool, pos or None)] for the active memories; mh / mh_kw: our MaskHeadSegLevel and
    the keyword tensors Query3DUnified binds to it (in-loop mask head).  Returns (q_out (B,N,D) fp32, predictions, saved)."""
    dev = query.device
    B, N, D = query.shape
    H, L = enc.num_heads, enc.num_layers
    R = B * N
    # weights change every step (and fused optimizers update them without bumping the tensors' version counters):
    # one pq3d_pack_segments launch rewrites every bf16 operand copy, plain and transposed, from the live parameters
    pk = enc.packed(dev, train=True)
    program = [g for g in enc._program() if len(g) > 0]
    sv: Dict = dict(B=B, N=N, D=D, H=H, L=L, R=R, program=program, pk=pk, layers=[], mems={})
    # train-time dropout (QueryEncoderLayer(dropout=0.1), query_encoder.py:97): counter RNG keyed by a per-forward device
    # seed, so the backward regenerates every mask and a CUDA-graph replay draws fresh ones
    p_drop = float(enc.train_dropout) if enc.training else 0.0
    p_mem = float(getattr(enc, "layer_memory_dropout", enc.memory_dropout)) if enc.training else 0.0
    seed = None
    if p_drop > 0.0 and N > 128:
        raise NotImplementedError("pq3d_b200 training path: dropout needs the fused attention backward (N <= 128 "
                                  "queries); set `encoder.train_dropout = 0.0` for more queries")
    if enc.num_blocks > 1 and N > 128:
        raise NotImplementedError("pq3d_b200 training path: num_blocks > 1 needs the fused attention backward (N <= 128)")
    if p_drop > 0.0 or (mh is not None and mh.training and float(mh.cls_head[3].p) > 0.0):
        if getattr(enc, "_drop_seed", None) is None or enc._drop_seed.device != dev:
            enc._drop_seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int32, device=dev)
        enc._drop_seed.add_(0x3C6EF35F)
        seed = enc._drop_seed.clone()
    sv.update(p_drop=p_drop, seed=seed)
    spatial = enc.spatial_selfattn

    ws: Dict = {}
    for name, feat, mask, pos in mems:
        multi = isinstance(feat, (list, tuple))          # multi-scale voxel memory: layer i attends feat[i] (:90-91)
        S = (feat[0] if multi else feat).shape[1]
        Sp = ops.pad64(S)
        st = _Mem()
        st.name, st.S, st.Sp, st.S_pitch, st.has_pos, st.multi = name, S, Sp, Sp, pos is not None, multi
        pos32 = None if pos is None else pos.detach().contiguous().float()
        st.K = _e((B * Sp, L * D), bf16, dev)
        st.Vt = _e((L * D, B * Sp), bf16, dev)

        def ingest(f):
            xv_ = _e((B * Sp, D), bf16, dev)
            xk_ = _e((B * Sp, D), bf16, dev) if pos is not None else xv_
            ops.ingest_memory(f.detach().contiguous().float(), pos32, xk_ if pos is not None else None, xv_, Sp)
            return xk_, xv_
        if multi:
            if len(feat) < L:
                raise ValueError(f"memory '{name}': {len(feat)} feature scales for {L} layers")
            st.xk, st.xv = [], []
            for l in range(L):
                xk_, xv_ = ingest(feat[l])
                st.xk.append(xk_)
                st.xv.append(xv_)
                sl = slice(l * D, (l + 1) * D)
                ops.linear(xk_, pk.wk[name][sl], st.K[:, sl], M=B * Sp, N=D, K=D, bias=pk.bk[name][sl], ldc=L * D)
                ops.linear(pk.wv[name][sl], xv_, st.Vt[sl], M=D, N=B * Sp, K=D, bias=pk.bv[name][sl], bias_along_m=True)
        else:
            st.xk, st.xv = ingest(feat)
            ops.linear(st.xk, pk.wk[name], st.K, M=B * Sp, N=L * D, K=D, bias=pk.bk[name])
            ops.linear(pk.wv[name], st.xv, st.Vt, M=L * D, N=B * Sp, K=D, bias=pk.bv[name], bias_along_m=True)
        enc._set_mask(st, mask, B, N, H, ws, dev)
        sv["mems"][name] = st

    q32 = query.detach().reshape(R, D).float().contiguous()
    qpos = query_pos.detach().reshape(R, D).float().contiguous()
    xq, xv = _e((R, D), bf16, dev), _e((R, D), bf16, dev)
    ops.cast_bf16(q32, xq, add=qpos)
    ops.cast_bf16(q32, xv)
    qbits = ops.pack_mask(query_masks.contiguous())
    sbias = pw = None
    if pk.loc_w is not None:
        if pairwise_locs is None:
            raise ValueError("spatial_selfattn=True needs pairwise_locs (B, N, N, 5)")
        pw = pairwise_locs.detach().float().contiguous()
        sbias = _e((L, B, H, N, ops.bias_ld(N)), f32, dev)
        ops.spatial_bias(pw, pk.loc_w, pk.loc_b, sbias)
    sv.update(qpos=qpos, qbits=qbits, sbias=sbias, pw=pw)
    Np = ops.pad8(N)

    def add_ln(y, res, w, G, site_id, row_w=None, **outs):
        if p_drop > 0.0 or row_w is not None:
            ops.add_layernorm_train(y, res, w["gamma"], w["beta"], w["eps"], R, D, G=G, y_group_stride=R * D,
                                    drop_p=p_drop, seed=seed, site=site_id, row_w=row_w, rows_per_scene=N, **outs)
        else:
            ops.add_layernorm(y, res, w["gamma"], w["beta"], w["eps"], R, D, G=G, y_group_stride=R * D, **outs)

    mht = None
    if mh is not None:
        from .mask_head import MaskHeadTrain
        mht = MaskHeadTrain(mh, mh_kw["seg_fts_for_match"], mh_kw["seg_masks"], B, N, seed)
    sv["mht"] = mht
    preds: List[torch.Tensor] = []
    for step in range(enc.num_blocks * L):
        i = step % L
        lw = pk.layers[i]
        lay = dict(groups=[], i=i, blk=step // L)
        step_masks: Dict = {}
        if mht is not None:
            # in-loop mask head on the layer's input query (query_encoder.py:78-81); its attention mask (detached, bool)
            # replaces the scene memories' masks for this layer when use_self_mask (:82-88)
            cls, logits, attn, lay["mh"] = mht.call(q32)
            preds += [cls, logits]
            if enc.use_self_mask:
                fixed = torch.empty(attn.shape, dtype=torch.bool, device=dev)
                tiles = torch.empty(B, dtype=torch.int32, device=dev)
                bits = ops.pack_mask(attn, None, unmask_full_rows=True, mask_fixed=fixed.view(torch.uint8),
                                     active_tiles=tiles)
                sv["final_mask"] = fixed
                for m in sv["mems"]:
                    if m != "prompt":
                        step_masks[m] = (bits, (bits.stride(0), 0, bits.stride(1)), tiles)
        elif enc.use_self_mask:
            raise ValueError("use_self_mask=True needs a mask_head that returns an attention mask")

        def mem_mask(m):
            st_ = sv["mems"][m]
            return step_masks.get(m, (st_.bits, st_.strides, st_.tiles))
        def ca_group(gi, grp, xq_in, res, **outs):
            """One cross-attention group (a parallel_ca / sequential_ca call) from the query operands xq_in / res:
            Q projection, one attention launch over the group's memories, grouped out-projection, add + LayerNorm
            (+ sublayer dropout, memory-dropout weights) into `outs`.  Returns the record the backward needs."""
            g = len(grp)
            w = lw["groups"][grp]
            Q = _e((R, g * D), bf16, dev)
            ops.linear(xq_in, w["wq"], Q, M=R, N=g * D, K=D, bias=w["bq"], alpha=ops.Q_SCALE, alpha_ncols=g * D)
            O = _e((g, R, D), bf16, dev)
            st_m, st_l = _e((g, B, H, N), f32, dev), _e((g, B, H, N), f32, dev)
            am = [ops.AttnMemory(sv["mems"][m].K, i * D, sv["mems"][m].Vt, i * D, sv["mems"][m].S, sv["mems"][m].Sp,
                                 mem_mask(m)[0], *mem_mask(m)[1], kv_tiles=mem_mask(m)[2]) for m in grp]
            psites = [rng.site(i, rng.SITE_CA_PROBS + enc.memories.index(m)) for m in grp]
            ops.attention(Q, D, am, O, R * D, B, H, N, True, stats=(st_m, st_l), drop_p=p_drop, seed=seed, sites=psites)
            y = _e((g, R, D), f32, dev)
            ops.linear(O.view(g * R, D), w["wo"], y, M=R, N=D, K=D, bias=w["bo"], bias_group_stride=D, groups=g,
                       a_group_rows=R, w_group_rows=D, ldc=D, c_group_stride=R * D)
            # train-time memory dropout of a parallel group (query_encoder.py:145-152): keep each (scene, memory) with
            # probability 1 - p, keep all when none survived, average over the survivors
            row_w = None
            if p_mem > 0.0 and g > 1 and enc._group_is_parallel(gi):
                keep = torch.rand(B, g, device=dev) > p_mem
                keep = torch.logical_or(keep, (keep.sum(dim=1) == 0).unsqueeze(-1))
                row_w = (keep.float() / keep.sum(dim=1, keepdim=True).float()).contiguous()
                enc.last_memory_keep.append(keep)
            add_ln(y, res, w, g, rng.site(i, rng.SITE_CA_SUBLAYER + gi), row_w, **outs)
            return dict(grp=grp, gi=gi, xq=xq_in, Q=Q, O=O, m=st_m, l=st_l, y=y, res=res, row_w=row_w,
                        masks={m: mem_mask(m) for m in grp})

        if enc.structure == "gate":
            # prompt = CA(query); gate = sigmoid(gate_proj(prompt)); update = parallel_ca(query, scene memories);
            # query = (1 - gate) * query + gate * update   (query_encoder.py:166-170) — both groups read the SAME query
            grp_p, grp_s = program
            pb16 = _e((R, D), bf16, dev)
            rec_p = ca_group(0, grp_p, xq, q32, out_bf16=pb16)
            gl = _e((R, D), f32, dev)
            ops.linear(pb16, lw["gate"]["w"], gl, M=R, N=D, K=D, bias=lw["gate"]["b"])
            upd = _e((R, D), f32, dev)
            rec_s = ca_group(1, grp_s, xq, q32, out_f32=upd)
            q_new, xq_new, xv_new = _e((R, D), f32, dev), _e((R, D), bf16, dev), _e((R, D), bf16, dev)
            ops.gate_mix(gl, q32, upd, q_new)
            ops.cast_bf16(q_new, xq_new, add=qpos)
            ops.cast_bf16(q_new, xv_new)
            lay["groups"] = [rec_p, rec_s]
            lay["gate"] = dict(gl=gl, upd=upd, q=q32, pb16=pb16)
            q32, xq, xv = q_new, xq_new, xv_new
        else:
            for gi, grp in enumerate(program):
                q_new, xq_new, xv_new = _e((R, D), f32, dev), _e((R, D), bf16, dev), _e((R, D), bf16, dev)
                lay["groups"].append(ca_group(gi, grp, xq, q32, pos=qpos, out_f32=q_new, out_bf16=xv_new,
                                              out_pos_bf16=xq_new))
                q32, xq, xv = q_new, xq_new, xv_new
        sa = lw["sa"]
        QK = _e((R, 2 * D), bf16, dev)
        ops.linear(xq, sa["wqk"], QK, M=R, N=2 * D, K=D, bias=sa["bqk"], alpha=ops.Q_SCALE, alpha_ncols=D)
        Vt = _z((D, B * Np), bf16, dev)
        ops.linear(sa["wv"], xv, Vt, M=D, N=N, K=D, bias=sa["bv"], bias_along_m=True, groups=B, a_group_rows=0,
                   w_group_rows=N, ldc=B * Np, c_group_stride=Np)
        Os = _e((1, R, D), bf16, dev)
        s_m, s_l = _e((1, B, H, N), f32, dev), _e((1, B, H, N), f32, dev)
        mem = ops.AttnMemory(QK, D, Vt, 0, N, N, qbits, qbits.stride(0), 0, 0, Vt_pitch=Np)
        # MultiHeadAttentionSpatial applies no dropout to its probabilities (transformers.py:158-240: the ctor's
        # `dropout` is unused); the stock MHA of SelfAttentionLayer does (query_encoder.py:195)
        sa_p = 0.0 if spatial else p_drop
        ops.attention(QK, 0, [mem], Os, R * D, B, H, N, False, None if sbias is None else sbias[i], stats=(s_m, s_l),
                      drop_p=sa_p, seed=seed, sites=[rng.site(i, rng.SITE_SA_PROBS)])
        ys = _e((R, D), f32, dev)
        ops.linear(Os.view(R, D), sa["wo"], ys, M=R, N=D, K=D, bias=sa["bo"])
        q_new, xv_new = _e((R, D), f32, dev), _e((R, D), bf16, dev)
        add_ln(ys, q32, sa, 1, rng.site(i, rng.SITE_SA_SUBLAYER), out_f32=q_new, out_bf16=xv_new)
        lay["sa"] = dict(xq=xq, xv=xv, QK=QK, Vt=Vt, O=Os, m=s_m, l=s_l, y=ys, res=q32)
        q32, xv = q_new, xv_new
        ff = lw["ffn"]
        F = ff["F"]
        h = _e((R, F), bf16, dev)
        ops.linear(xv, ff["w1"], h, M=R, N=F, K=D, bias=ff["b1"], relu=True)
        if p_drop > 0.0:                 # dropout(relu(linear1(x))) (query_encoder.py:384); h is saved dropped + rescaled
            ops.dropout_bf16(h, p_drop, seed, rng.site(i, rng.SITE_FFN_HIDDEN))
        yf = _e((R, D), f32, dev)
        ops.linear(h, ff["w2"], yf, M=R, N=D, K=F, bias=ff["b2"])
        q_new, xq_new, xv_new = _e((R, D), f32, dev), _e((R, D), bf16, dev), _e((R, D), bf16, dev)
        add_ln(yf, q32, ff, 1, rng.site(i, rng.SITE_FFN_SUBLAYER), pos=qpos, out_f32=q_new, out_bf16=xv_new,
               out_pos_bf16=xq_new)
        lay["ffn"] = dict(xv=xv, h=h, y=yf, res=q32)
        q32, xq, xv = q_new, xq_new, xv_new
        sv["layers"].append(lay)
    return q32.view(B, N, D), preds, sv


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
class _Bwd:
    """One backward pass.  Stream plan: the MAIN stream carries the critical chain (LayerNorm backward -> dgrad ->
    attention backward -> dgrad ...); everything that only feeds parameter gradients (column sums, operand
    transposes, wgrad GEMMs) is forked onto a SIDE stream as soon as its inputs exist, and the independent
    per-memory attention backwards of a parallel group each take their own stream.  All streams join before the
    gradients are handed to autograd.  Every tensor that crosses streams is kept alive in `self.keep` until the
    join (the caching allocator reuses a freed block on its allocation stream immediately)."""

    def __init__(self, enc, sv):
        self.enc, self.sv = enc, sv
        self.pk = sv["pk"]
        self.B, self.N, self.D, self.H, self.L, self.R = (sv[k] for k in ("B", "N", "D", "H", "L", "R"))
        self.Rp = ops.pad64(self.R)
        self.grads: Dict[str, torch.Tensor] = {}
        self.pending: Dict[str, List[torch.Tensor]] = {}
        self.sent = set()              # gradients already handed to the multi-GPU bucket all-reduce
        self.dev = sv["qpos"].device
        self.keep: List = []
        self.streams_on = bool(enc.train_streams)    # False: everything on the caller's stream, in program order
        self.main = torch.cuda.current_stream(self.dev) if self.streams_on else None
        n_par = max([len(g) for g in sv["program"]] + [1])
        pool = _side_streams(self.dev, n_par) if self.streams_on else []
        self.side = pool[0] if pool else None
        self.par = pool[1:] if pool else []
        self.forked = set()
        n_mem = len(enc.memories)
        # in_proj gradients of every cross-attention are written in place by the wgrad GEMMs: rows [0, D) from the
        # query side, [D, 2D) / [2D, 3D) from the hoisted K / V projections.  One buffer per block (the same layers are
        # re-applied num_blocks times): block 0 also receives the K / V rows, the others start from zero.
        self.G_blk = [(_e if blk == 0 else _z)((self.L, n_mem, 3 * self.D, self.D), f32, self.dev)
                      for blk in range(enc.num_blocks)]
        self.G_ca = self.G_blk[0]
        self.kv_written = set()

    # ---- streams ----------------------------------------------------------------------------------
    def on(self, stream, *deps):
        """Context: run on `stream` after everything issued so far on the main stream; deps are kept alive."""
        self.keep.extend(d for d in deps if d is not None)
        if not self.streams_on:
            return contextlib.nullcontext()
        if stream is None:
            return torch.cuda.stream(self.main)
        stream.wait_stream(torch.cuda.current_stream(self.dev))
        self.forked.add(stream)
        return torch.cuda.stream(stream)

    def join(self, streams=None):
        if not self.streams_on:
            return
        cur = torch.cuda.current_stream(self.dev)
        for st in (list(self.forked) if streams is None else streams):
            if st is not None and st in self.forked:
                cur.wait_stream(st)
                self.forked.discard(st)

    # ---- small helpers --------------------------------------------------------------------------
    def acc(self, name: str, g: torch.Tensor):
        """Record a gradient contribution.  Contributions are produced on different streams (main, side, per-memory);
        they are only SUMMED in `finish_grads`, after every stream has joined — an add issued here would read tensors
        whose producer may still be queued on another stream (num_blocks > 1, share_layer)."""
        name = self.sv["canon"].get(name, name)           # share_layer: one parameter under several names
        self.pending.setdefault(name, []).append(g)

    def finish_grads(self):
        """After the final join: one tensor per parameter (sums where a parameter received several contributions)."""
        for name, parts in self.pending.items():
            g = parts[0]
            for extra in parts[1:]:
                g = g + extra
            self.grads[name] = g
        self.pending.clear()

    def tcast(self, x, rows, cols, want_c=False, gate=None, scale=1.0):
        """x [rows, cols] (fp32/bf16) -> (bf16 x^T [cols, pad64(rows)], bf16 copy or None)."""
        xt = _e((cols, ops.pad64(rows)), bf16, self.dev)
        xc = _e((rows, cols), bf16, self.dev) if want_c else None
        ops.transpose_cast(x, xt, xc, gate=gate, scale=scale)
        return xt, xc

    def drop(self, i, kind):
        """kwargs of the dropout stream `kind` of layer i for layernorm_bwd / attention_bwd."""
        return dict(drop_p=self.sv["p_drop"], seed=self.sv["seed"], site=rng.site(i, kind))

    def wgrad(self, dyT, xT, n_out, n_in, out=None, groups=1, c_group_stride=0):
        """dW [n_out, n_in] = dy^T x from the K-major transposes (contraction over padded rows).  With groups > 1,
        dyT holds `groups` row blocks of n_out rows and block g lands at out + g*c_group_stride."""
        dW = _e((n_out, n_in), f32, self.dev) if out is None else out
        ops.linear(dyT, xT, dW, M=n_out, N=n_in, K=dyT.shape[1], groups=groups, a_group_rows=n_out if groups > 1 else 0,
                   w_group_rows=0, ldc=n_in, c_group_stride=c_group_stride)
        return dW

    def dgrad(self, dy16, w_t, n_in, out_dtype=f32):
        """d_x [rows, n_in] = dy W, W^T given as [n_in, n_out] bf16."""
        rows, n_out = dy16.shape
        dx = _e((rows, n_in), out_dtype, self.dev)
        ops.linear(dy16, w_t, dx, M=rows, N=n_in, K=n_out)
        return dx

    def colsum(self, x, gate=None, scale=1.0):
        out = _e((x.shape[1],), f32, self.dev)
        ops.colsum(x, out, gate=gate, scale=scale)
        return out

    # ---- attention backward for one (scene batch, key set) -----------------------------------------
    def attention_bwd(self, Qv, Kv, Vv, Ktp, dO2d, O2d, st_m, st_l, S, ld, dQv, dKv, dVv, bias=None, mask_bits=None,
                      mask_strides=(0, 0, 0)):
        """Qv (B,H,N,64) log2-domain queries; Kv, Vv (B,H,S,64); Ktp (B,H,64,ld) K^T with zero/finite pad;
        dO2d, O2d [R, D]; outputs written through the strided views dQv (scaled by Q_SCALE), dKv, dVv.
        Returns dS (bf16 [B,H,N,ld]) for the spatial-bias backward."""
        B, H, N, dev = self.B, self.H, self.N, self.dev
        Npad = ops.pad64(N)
        S2 = _e((B, H, N, ld), f32, dev)
        dP = _e((B, H, N, ld), f32, dev)
        sub = lambda t: t.as_strided((B, H, N, S), t.stride())     # noqa: E731
        ops.bgemm(Qv, Kv, sub(S2))
        dOv = _heads(dO2d, B, N, N, H)
        ops.bgemm(dOv, Vv, sub(dP))
        delta = _e((B, H, N), f32, dev)
        ops.attn_delta(dO2d, O2d, delta, B, H, N)
        dS = _e((B, H, N, ld), bf16, dev)
        Pt, dSt = _e((B, H, ld, Npad), bf16, dev), _e((B, H, ld, Npad), bf16, dev)
        ops.softmax_bwd(S2, dP, delta, st_m, st_l, None, dS, Pt, dSt, B, H, N, S, ld, Npad, bias=bias, mask_bits=mask_bits,
                        mask_strides=mask_strides)
        dOt, Qt = _e((B, H, 64, Npad), bf16, dev), _e((B, H, 64, Npad), bf16, dev)
        ops.transpose_cast(dOv, dOt)
        ops.transpose_cast(Qv, Qt)
        ops.bgemm(Pt[:, :, :S], dOt, dVv)
        ops.bgemm(dSt[:, :, :S], Qt, dKv)
        ops.bgemm(dS, Ktp, dQv, alpha=ops.Q_SCALE)
        return dS

    def fused_ok(self) -> bool:
        return self.N <= 128 and getattr(self.enc, "fused_attn_bwd", True)

    def attention_bwd_fused(self, Q, q_col0, dO, O2d, K, k_col0, V, v_col0, S, S_pitch, st_m, st_l, dK, dk_col0, dV,
                            dv_col0, dQ32, dq_col0, bias=None, mask_bits=None, mask_strides=(0, 0, 0), dS_out=None,
                            drop=None):
        """One pq3d_attention_bwd launch (+ the row-dot delta): scores are recomputed and consumed on chip."""
        B, H, N = self.B, self.H, self.N
        delta = _e((B, H, N), f32, self.dev)
        ops.attn_delta(dO, O2d, delta, B, H, N)
        ops.attention_bwd(Q, q_col0, dO, 0, K, k_col0, V, v_col0, S, S_pitch, st_m, st_l, delta, dK, dk_col0, dV, dv_col0,
                          dQ32, dq_col0, B, H, N, mask_bits=mask_bits, mask_strides=mask_strides, bias=bias, dS_out=dS_out,
                          **(drop or {}))

    # ---- blocks ---------------------------------------------------------------------------------------
    def ffn_bwd(self, i, s, d_out):
        R, D, dev, pk = self.R, self.D, self.dev, self.pk
        ff = pk.layers[i]["ffn"]
        F = ff["F"]
        pre = f"unified_encoder.{i}.ffn."
        d_y = _e((R, D), f32, dev)
        dg, db = _z((1, D), f32, dev), _z((1, D), f32, dev)
        d_y16 = _e((R, D), bf16, dev)
        p_drop = self.sv["p_drop"]
        kscale = 1.0 / (1.0 - p_drop)
        d_res = _e((R, D), f32, dev) if p_drop > 0.0 else d_y        # without dropout the two gradients coincide
        ops.layernorm_bwd(s["y"], s["res"], ff["gamma"], d_out, ff["eps"], R, D, d_x=d_y, d_gamma=dg, d_beta=db, d_x16=d_y16,
                          d_res=d_res if p_drop > 0.0 else None, **self.drop(i, rng.SITE_FFN_SUBLAYER))
        self.acc(pre + "norm.weight", dg[0]); self.acc(pre + "norm.bias", db[0])
        with self.on(self.side, d_y, dg, db):
            self.acc(pre + "linear2.bias", self.colsum(d_y))
            d_yT, _ = self.tcast(d_y, R, D)
            hT, _ = self.tcast(s["h"], R, F)
            self.acc(pre + "linear2.weight", self.wgrad(d_yT, hT, D, F))
        d_h = self.dgrad(d_y16, pk.T(ff["w2"]), F)
        d_preT, d_pre16 = self.tcast(d_h, R, F, want_c=True, gate=s["h"], scale=kscale)   # relu and hidden-dropout gates
        with self.on(self.side, d_h, d_preT, d_y16):
            self.acc(pre + "linear1.bias", self.colsum(d_h, gate=s["h"], scale=kscale))
            xvT, _ = self.tcast(s["xv"], R, D)
            self.acc(pre + "linear1.weight", self.wgrad(d_preT, xvT, F, D))
        d_xv = self.dgrad(d_pre16, pk.T(ff["w1"]), D)
        d_in = _e((R, D), f32, dev)
        ops.add3(d_res, d_xv, None, d_in)
        self.keep += [d_pre16, d_xv, d_res]
        return d_in

    def sa_bwd(self, i, s, d_out, d_pos):
        R, D, B, N, H, dev, pk = self.R, self.D, self.B, self.N, self.H, self.dev, self.pk
        sa = pk.layers[i]["sa"]
        spatial = sa["loc_w"] is not None
        pre = f"unified_encoder.{i}.self_attn."
        a = pre + "self_attn."
        d_y = _e((R, D), f32, dev)
        dg, db = _z((1, D), f32, dev), _z((1, D), f32, dev)
        d_y16 = _e((R, D), bf16, dev)
        p_drop = self.sv["p_drop"]
        d_res = _e((R, D), f32, dev) if p_drop > 0.0 else d_y
        ops.layernorm_bwd(s["y"], s["res"], sa["gamma"], d_out, sa["eps"], R, D, d_x=d_y, d_gamma=dg, d_beta=db, d_x16=d_y16,
                          d_res=d_res if p_drop > 0.0 else None, **self.drop(i, rng.SITE_SA_SUBLAYER))
        self.acc(pre + "norm.weight", dg[0]); self.acc(pre + "norm.bias", db[0])
        O2d = s["O"].view(R, D)
        if not spatial:
            g_in, g_inb = _e((3 * D, D), f32, dev), _e((3 * D,), f32, dev)
        with self.on(self.side, d_y, dg, db):
            d_bo = self.colsum(d_y)
            d_yT, _ = self.tcast(d_y, R, D)
            OT, _ = self.tcast(O2d, R, D)
            d_wo = self.wgrad(d_yT, OT, D, D)
            self.acc(a + ("fc.weight" if spatial else "out_proj.weight"), d_wo)
            self.acc(a + ("fc.bias" if spatial else "out_proj.bias"), d_bo)
        dO = self.dgrad(d_y16, pk.T(sa["wo"]), D, out_dtype=bf16)
        # operands of the attention backward
        QK = s["QK"]
        Np8, ld = ops.pad8(N), ops.pad64(N)
        Qv, Kv = _heads(QK, B, N, N, H, 0), _heads(QK, B, N, N, H, D)
        V = _e((R, D), bf16, dev)                                       # V back in row-major from V^T [D, B*Np8]
        Vt = s["Vt"]
        ops.transpose_cast(Vt.as_strided((B, 1, D, N), (Np8, 0, Vt.stride(0), 1)), V.view(B, 1, N, D))
        Vv = _heads(V, B, N, N, H)
        dQK = _e((R, 2 * D), bf16, dev)
        dV = _e((R, D), bf16, dev)
        qbits = self.sv["qbits"]
        bias = None if self.sv["sbias"] is None else self.sv["sbias"][i]
        Ktp = None
        if self.fused_ok():
            dQ32 = _z((R, D), f32, dev)
            dS = _e((B, H, N, ld), bf16, dev) if spatial else None
            self.attention_bwd_fused(QK, 0, dO, O2d, QK, D, V, 0, N, N, s["m"][0], s["l"][0], dQK, D, dV, 0, dQ32, 0,
                                     bias=bias, mask_bits=qbits, mask_strides=(qbits.stride(0), 0, 0), dS_out=dS,
                                     drop=None if (spatial or p_drop == 0.0) else self.drop(i, rng.SITE_SA_PROBS))
            ops.transpose_cast(dQ32, None, dQK[:, :D])
            self.keep.append(dQ32)
        else:
            Ktp = _e((B, H, 64, ld), bf16, dev)
            ops.transpose_cast(Kv, Ktp)
            dS = self.attention_bwd(Qv, Kv, Vv, Ktp, dO, O2d, s["m"][0], s["l"][0], N, ld, _heads(dQK, B, N, N, H, 0),
                                    _heads(dQK, B, N, N, H, D), _heads(dV, B, N, N, H), bias=bias, mask_bits=qbits,
                                    mask_strides=(qbits.stride(0), 0, 0))
        with self.on(self.side, dQK, dV, dS, d_y16, dO):
            d_bqk = self.colsum(dQK)
            dQKT, _ = self.tcast(dQK, R, 2 * D)
            xqT, _ = self.tcast(s["xq"], R, D)
            d_bv = self.colsum(dV)
            dVT, _ = self.tcast(dV, R, D)
            xvT, _ = self.tcast(s["xv"], R, D)
            if spatial:
                d_wqk = self.wgrad(dQKT, xqT, 2 * D, D)
                d_wv = self.wgrad(dVT, xvT, D, D)
                self.acc(a + "w_qs.weight", d_wqk[:D]); self.acc(a + "w_ks.weight", d_wqk[D:])
                self.acc(a + "w_qs.bias", d_bqk[:D]); self.acc(a + "w_ks.bias", d_bqk[D:])
                self.acc(a + "w_vs.weight", d_wv); self.acc(a + "w_vs.bias", d_bv)
                d_lw, d_lb = _z((H, 5), f32, dev), _z((H,), f32, dev)
                ops.spatial_bias_bwd(self.sv["pw"], sa["loc_w"], sa["loc_b"], dS, ld, d_lw, d_lb, B, H, N)
                self.acc(a + "pairwise_loc_fc.weight", d_lw); self.acc(a + "pairwise_loc_fc.bias", d_lb)
            else:
                self.wgrad(dQKT, xqT, 2 * D, D, out=g_in[:2 * D])
                self.wgrad(dVT, xvT, D, D, out=g_in[2 * D:])
                g_inb[:2 * D].copy_(d_bqk); g_inb[2 * D:].copy_(d_bv)
                self.acc(a + "in_proj_weight", g_in); self.acc(a + "in_proj_bias", g_inb)
        d_xq = self.dgrad(dQK, pk.T(sa["wqk"]), D)
        d_xv = self.dgrad(dV, pk.T(sa["wv"]), D)
        d_in = _e((R, D), f32, dev)
        ops.add3(d_res, d_xq, d_xv, d_in)
        self.keep.append(d_res)
        with self.on(self.side, d_xq):
            ops.add3(d_pos, d_xq, None, d_pos)
        self.keep += [d_xq, d_xv, V, Ktp]
        return d_in

    def group_bwd(self, i, s, d_out, d_pos, mem_grads, blk=0):
        R, D, B, N, H, dev, pk = self.R, self.D, self.B, self.N, self.H, self.dev, self.pk
        grp = s["grp"]
        g = len(grp)
        w = pk.layers[i]["groups"][grp]
        idx = [self.enc.memories.index(m) for m in grp]
        d_y = _e((g, R, D), f32, dev)
        d_res = _e((R, D), f32, dev)
        dg, db = _z((g, D), f32, dev), _z((g, D), f32, dev)
        d_y16 = _e((g, R, D), bf16, dev)
        ops.layernorm_bwd(s["y"], s["res"], w["gamma"], d_out, w["eps"], R, D, G=g, y_group_stride=R * D, d_x=d_y,
                          dx_group_stride=R * D, d_res=d_res, d_gamma=dg, d_beta=db, d_x16=d_y16, row_w=s["row_w"],
                          rows_per_scene=N, **self.drop(i, rng.SITE_CA_SUBLAYER + s["gi"]))
        fused = self.fused_ok()
        dQ = _z((R, g * D), f32, dev) if fused else _e((R, g * D), bf16, dev)     # fused: fp32, accumulated by atomics
        self.keep += [d_y, dg, db, dQ, d_out]
        self.keep.append(d_y16)
        with self.on(self.side):
            for jj, j in enumerate(idx):
                pre = f"unified_encoder.{i}.cross_attn_list.{j}."
                self.acc(pre + "norm.weight", dg[jj]); self.acc(pre + "norm.bias", db[jj])
                self.acc(pre + "multihead_attn.out_proj.bias", self.colsum(d_y[jj]))
                d_yT, _ = self.tcast(d_y[jj], R, D)
                OT, _ = self.tcast(s["O"][jj], R, D)
                self.acc(pre + "multihead_attn.out_proj.weight", self.wgrad(d_yT, OT, D, D))
        used = []
        for jj, m in enumerate(grp):
            # each memory's chain (dgrad -> attention backward) is independent of the others': own stream
            stream = self.par[jj - 1] if (jj > 0 and jj - 1 < len(self.par)) else None
            st = self.sv["mems"][m]
            with self.on(stream):
                used.append(stream)
                dO = self.dgrad(d_y16[jj], pk.T(w["wo"], jj * D, D), D, out_dtype=bf16)
                mg = mem_grads[m]
                S, Sp = st.S, st.Sp
                m_bits, m_strides, _ = s["masks"][m]
                if fused:
                    first = (m, i) not in self.kv_written          # later blocks of the same layer: accumulate
                    self.kv_written.add((m, i))
                    if first:
                        dK_dst, dV_dst, col = mg["dK"], mg["dV"], i * D
                    else:
                        dK_dst, dV_dst, col = _z((B * Sp, D), bf16, dev), _z((B * Sp, D), bf16, dev), 0
                    self.attention_bwd_fused(s["Q"], jj * D, dO, s["O"][jj], st.K, i * D, mg["V"], i * D, S, Sp, s["m"][jj],
                                             s["l"][jj], dK_dst, col, dV_dst, col, dQ, jj * D, mask_bits=m_bits,
                                             mask_strides=m_strides,
                                             drop=None if self.sv["p_drop"] == 0.0 else
                                             self.drop(i, rng.SITE_CA_PROBS + self.enc.memories.index(m)))
                    if not first:
                        mg["dK"][:, i * D:(i + 1) * D].add_(dK_dst)
                        mg["dV"][:, i * D:(i + 1) * D].add_(dV_dst)
                else:
                    Qv = _heads(s["Q"], B, N, N, H, jj * D)
                    Kv = _heads(st.K, B, S, Sp, H, i * D)
                    Vv = _heads(mg["V"], B, S, Sp, H, i * D)
                    Kt = mg["Kt"]                                               # [L*D, B*Sp]
                    Ktp = Kt.as_strided((B, H, 64, Sp), (Sp, 64 * Kt.stride(0), Kt.stride(0), 1), i * D * Kt.stride(0))
                    self.attention_bwd(Qv, Kv, Vv, Ktp, dO, s["O"][jj], s["m"][jj], s["l"][jj], S, Sp,
                                       _heads(dQ, B, N, N, H, jj * D), _heads(mg["dK"], B, S, Sp, H, i * D),
                                       _heads(mg["dV"], B, S, Sp, H, i * D), mask_bits=m_bits, mask_strides=m_strides)
                self.keep.append(dO)
        self.join([st_ for st_ in used if st_ is not None])
        G = self.G_blk[blk][i]
        step = idx[1] - idx[0] if g > 1 else 0
        regular = all(idx[k + 1] - idx[k] == step for k in range(g - 1)) and (g == 1 or step > 0)
        dQT, dQ16 = self.tcast(dQ, R, g * D, want_c=fused)
        if not fused:
            dQ16 = dQ
        self.keep += [dQT, dQ16]
        with self.on(self.side):
            d_bq = self.colsum(dQ)
            xqT, _ = self.tcast(s["xq"], R, D)
            if regular:
                self.wgrad(dQT, xqT, D, D, out=G[idx[0], :D], groups=g, c_group_stride=step * 3 * D * D)
            else:
                for jj, j in enumerate(idx):
                    self.wgrad(dQT[jj * D:(jj + 1) * D], xqT, D, D, out=G[j, :D])
        mg_q = mem_grads["_q"]
        for jj, j in enumerate(idx):          # d_bq lives on the side stream: summed there (memory-side tail of run())
            mg_q.setdefault((i, j), []).append(d_bq[jj * D:(jj + 1) * D])
        d_xq = self.dgrad(dQ16, pk.T(w["wq"]), D)
        d_in = _e((R, D), f32, dev)
        ops.add3(d_res, d_xq, None, d_in)
        with self.on(self.side, d_xq):
            ops.add3(d_pos, d_xq, None, d_pos)
        self.keep += [d_xq, d_res]
        return d_in

    def gate_bwd(self, i, lay, d_out, d_pos, mem_grads):
        """structure 'gate': through the mix, gate_proj, and BOTH cross-attention groups (same input query)."""
        R, D, dev, pk = self.R, self.D, self.dev, self.pk
        gt = lay["gate"]
        wg = pk.layers[i]["gate"]["w"]
        d_gl, d_gl16 = _e((R, D), f32, dev), _e((R, D), bf16, dev)
        d_upd, d_qdir = _e((R, D), f32, dev), _e((R, D), f32, dev)
        ops.gate_mix_bwd(gt["gl"], gt["q"], gt["upd"], d_out, d_gl, d_gl16, d_upd, d_qdir)
        with self.on(self.side, d_gl, d_out):
            self.acc(f"unified_encoder.{i}.gate_proj.bias", self.colsum(d_gl))
            d_glT, _ = self.tcast(d_gl, R, D)
            pbT, _ = self.tcast(gt["pb16"], R, D)
            self.acc(f"unified_encoder.{i}.gate_proj.weight", self.wgrad(d_glT, pbT, D, D))
        d_pf = self.dgrad(d_gl16, pk.T(wg), D)
        rec_p, rec_s = lay["groups"]
        d_in_p = self.group_bwd(i, rec_p, d_pf, d_pos, mem_grads, lay["blk"])
        d_in_s = self.group_bwd(i, rec_s, d_upd, d_pos, mem_grads, lay["blk"])
        d_in = _e((R, D), f32, dev)
        ops.add3(d_qdir, d_in_p, d_in_s, d_in)
        self.keep += [d_gl16, d_upd, d_qdir, d_pf, d_in_p, d_in_s]
        return d_in

    # ---- whole decoder ------------------------------------------------------------------------------------
    def run(self, d_out: torch.Tensor, d_preds=()):
        sv, B, N, D, L, R, dev, pk = self.sv, self.B, self.N, self.D, self.L, self.R, self.dev, self.pk
        n_mem = len(self.enc.memories)
        mem_grads: Dict = {"_q": {}}
        for name, st in sv["mems"].items():
            V = _e((B * st.Sp, L * D), bf16, dev)
            ops.transpose_cast(st.Vt, V)                               # V^T [L*D, B*Sp] -> V [B*Sp, L*D]
            Kt = None
            if not self.fused_ok():
                Kt = _e((L * D, B * st.Sp), bf16, dev)
                ops.transpose_cast(st.K, Kt)
            mem_grads[name] = dict(V=V, Kt=Kt, dK=_z((B * st.Sp, L * D), bf16, dev), dV=_z((B * st.Sp, L * D), bf16, dev))
        d_q = d_out.detach().reshape(R, D).float().contiguous()
        d_pos = _z((R, D), f32, dev)
        mht = sv.get("mht")
        # multi-GPU: gradient buckets leave for the all-reduce while the backward is still running (dist.FlatGradAllReduce)
        sink = getattr(self.enc, "grad_sink", None)
        if sink is not None and not (sink.enabled and dist_ready()):
            sink = None
        if sink is not None:
            sink.begin_backward()
        for step in reversed(range(len(sv["layers"]))):
            lay = sv["layers"][step]
            i = lay["i"]
            d_q = self.ffn_bwd(i, lay["ffn"], d_q)
            d_q = self.sa_bwd(i, lay["sa"], d_q, d_pos)
            if "gate" in lay:
                d_q = self.gate_bwd(i, lay, d_q, d_pos, mem_grads)
            else:
                for s in reversed(lay["groups"]):
                    d_q = self.group_bwd(i, s, d_q, d_pos, mem_grads, lay["blk"])
            if mht is not None:                 # the mask head read this layer's input query
                d_cls, d_logits = d_preds[2 * step], d_preds[2 * step + 1]
                if d_cls is not None or d_logits is not None:
                    d_mh = mht.call_bwd(lay["mh"], d_cls, d_logits)
                    if d_mh is not None:
                        d_sum = _e((R, D), f32, dev)
                        ops.add3(d_q, d_mh, None, d_sum)
                        self.keep += [d_q, d_mh]
                        d_q = d_sum
            self.keep.append(d_q)
            if sink is not None:
                # layer i's query-side gradients are final (num_blocks == 1, no shared layers: one contribution each);
                # the in-projection rows of the K / V projections only complete in the memory-side tail below
                pre = f"unified_encoder.{i}."
                ready = {n: parts[0] for n, parts in self.pending.items()
                         if n.startswith(pre) and "in_proj" not in n and len(parts) == 1 and n not in self.sent}
                self.sent.update(ready)
                sink.reduce_async(("layer", i), ready, producers=[self.side] + list(self.par))
        # memory side: weight / bias gradients of the hoisted K and V projections, input gradients
        d_mem = {}
        G_b = _e((L, n_mem, 3 * D), f32, dev)
        for name, st in sv["mems"].items():
            mg = mem_grads[name]
            j = self.enc.memories.index(name)
            rows = B * st.Sp
            with self.on(self.side):
                d_bk, d_bv = self.colsum(mg["dK"]), self.colsum(mg["dV"])
                if st.multi:                     # one feature table per layer: a wgrad per layer
                    for l in range(L):
                        sl = slice(l * D, (l + 1) * D)
                        dKT, _ = self.tcast(mg["dK"][:, sl], rows, D)
                        xkT, _ = self.tcast(st.xk[l], rows, D)
                        self.wgrad(dKT, xkT, D, D, out=self.G_ca[l, j, D:2 * D])
                        dVT, _ = self.tcast(mg["dV"][:, sl], rows, D)
                        xvT = xkT if not st.has_pos else self.tcast(st.xv[l], rows, D)[0]
                        self.wgrad(dVT, xvT, D, D, out=self.G_ca[l, j, 2 * D:])
                else:
                    dKT, _ = self.tcast(mg["dK"], rows, L * D)
                    xkT, _ = self.tcast(st.xk, rows, D)
                    self.wgrad(dKT, xkT, D, D, out=self.G_ca[0, j, D:2 * D], groups=L, c_group_stride=n_mem * 3 * D * D)
                    dVT, _ = self.tcast(mg["dV"], rows, L * D)
                    xvT = xkT if not st.has_pos else self.tcast(st.xv, rows, D)[0]
                    self.wgrad(dVT, xvT, D, D, out=self.G_ca[0, j, 2 * D:], groups=L, c_group_stride=n_mem * 3 * D * D)
                G_b[:, j, D:2 * D].copy_(d_bk.view(L, D))
                G_b[:, j, 2 * D:].copy_(d_bv.view(L, D))
                for i in range(L):
                    parts = mem_grads["_q"][(i, j)]
                    G_b[i, j, :D].copy_(parts[0])
                    for extra in parts[1:]:
                        G_b[i, j, :D].add_(extra)
            for i in range(L):
                pre = f"unified_encoder.{i}.cross_attn_list.{j}.multihead_attn."
                for G in self.G_blk:
                    self.acc(pre + "in_proj_weight", G[i, j])
                self.acc(pre + "in_proj_bias", G_b[i, j])
            if sink is not None:
                # this memory's in-projection gradients (all layers) are complete once its wgrads above have run on the
                # side stream: send them now, the next memory's wgrad / dgrad work hides the transfer
                tag = f".cross_attn_list.{j}.multihead_attn.in_proj"
                ready = {n: parts[0] for n, parts in self.pending.items()
                         if tag in n and len(parts) == 1 and n not in self.sent}
                self.sent.update(ready)
                sink.reduce_async(("mem", j), ready, producers=[self.side])
            cut = lambda t: t.view(B, st.Sp, D)[:, :st.S]              # noqa: E731
            if st.multi:
                d_feats, d_p = [], None
                for l in range(L):
                    sl = slice(l * D, (l + 1) * D)
                    d_xk = self.dgrad(mg["dK"][:, sl], pk.T_mem("k", j, L, D)[:, sl], D)
                    d_xv = self.dgrad(mg["dV"][:, sl], pk.T_mem("v", j, L, D)[:, sl], D)
                    d_f = _e((rows, D), f32, dev)
                    ops.add3(d_xk, d_xv, None, d_f)
                    d_feats.append(cut(d_f))
                    if st.has_pos:
                        if d_p is None:
                            d_p = d_xk
                        else:
                            d_new = _e((rows, D), f32, dev)
                            ops.add3(d_p, d_xk, None, d_new)
                            self.keep.append(d_p)
                            d_p = d_new
                    self.keep += [d_xk, d_xv]
                d_mem[name] = (d_feats, cut(d_p) if st.has_pos else None)
            else:
                d_xk = self.dgrad(mg["dK"], pk.T_mem("k", j, L, D), D)
                d_xv = self.dgrad(mg["dV"], pk.T_mem("v", j, L, D), D)
                d_feat = _e((rows, D), f32, dev)
                ops.add3(d_xk, d_xv, None, d_feat)
                self.keep += [d_xk, d_xv]
                d_mem[name] = (cut(d_feat), cut(d_xk) if st.has_pos else None)
        mh_out = (None, None)
        if mht is not None:
            mh_out = mht.finish()
        self.join()
        self.finish_grads()
        if sink is not None:
            rest = {n: g for n, g in self.grads.items() if n not in self.sent}
            sink.reduce_async(("tail",), rest)
            sink.finish_backward(self.grads)
        self.keep.clear()
        return d_q.view(B, N, D), d_pos.view(B, N, D), d_mem, self.grads, mh_out


class DecoderFunction(torch.autograd.Function):
    """forward(enc, meta, query, query_pos, feat_0, pos_0, ..., *mask-head features, *decoder params, *mask-head params)
    -> (decoded queries (B, N, D), cls_0, mask_logits_0, cls_1, ...)."""

    @staticmethod
    def forward(ctx, enc, meta, query, query_pos, *rest):
        mems, o = [], 0
        for k, name in enumerate(meta["names"]):
            nf = meta["nfeat"][k]                          # 1, or one feature tensor per layer (multi-scale voxels)
            feat = rest[o] if nf == 1 and not meta["is_list"][k] else list(rest[o:o + nf])
            mems.append((name, feat, meta["masks"][k], rest[o + nf]))
            o += nf + 1
        ctx.n_mem_inputs = o
        out, preds, sv = forward(enc, query, query_pos, meta["query_masks"], mems, meta["pairwise_locs"], meta["mh"],
                                 meta["mh_kw"])
        sv["canon"] = meta["canon"]
        meta["final_mask"] = sv.get("final_mask")
        ctx.enc, ctx.sv, ctx.meta = enc, sv, meta
        ctx.in_dtypes = (query.dtype, query_pos.dtype, [None if t is None else t.dtype for t in rest[:o]])
        return (out, *preds)

    @staticmethod
    def backward(ctx, d_out, *d_preds):
        enc, sv, meta = ctx.enc, ctx.sv, ctx.meta
        if d_out is None:
            d_out = torch.zeros(sv["B"], sv["N"], sv["D"], dtype=f32, device=sv["qpos"].device)
        d_q, d_pos, d_mem, grads, (mh_grads, mh_dfeats) = _Bwd(enc, sv).run(d_out, d_preds)
        ctx.sv = None
        out = [None, None, d_q.to(ctx.in_dtypes[0]), d_pos.to(ctx.in_dtypes[1])]
        o = 0
        for k, name in enumerate(meta["names"]):
            d_feat, d_p = d_mem[name]
            for g in (d_feat if isinstance(d_feat, list) else [d_feat]):
                out.append(g.to(ctx.in_dtypes[2][o]) if ctx.needs_input_grad[4 + o] else None)
                o += 1
            out.append(d_p.to(ctx.in_dtypes[2][o]) if d_p is not None and ctx.needs_input_grad[4 + o] else None)
            o += 1
        for j in range(meta["n_mh_feats"]):
            g = None if mh_dfeats is None else mh_dfeats[j]
            out.append(g if (g is not None and ctx.needs_input_grad[4 + o + j]) else None)
        for pname, p in meta["param_names"]:
            g = grads.get(pname)
            out.append(None if g is None else g.reshape(p.shape).to(p.dtype))
        for pname, p in meta["mh_param_names"]:
            g = None if mh_grads is None else mh_grads.get(pname)
            out.append(None if g is None else g.reshape(p.shape).to(p.dtype))
        return tuple(out)


def run(enc, input_dict: dict, pairwise_locs, mask_head=None):
    """Entry used by QueryMaskEncoder.forward in training.  Returns (query, predictions_class, predictions_mask)."""
    mh, mh_kw = None, None
    if mask_head is not None:
        mh = enc._own_mask_head(mask_head)
        if mh is None:
            raise NotImplementedError(
                "pq3d_b200 training path: the in-loop mask head must be pq3d_b200.MaskHeadSegLevel bound with "
                "functools.partial(seg_fts_for_match=..., seg_masks=..., offline_attn_masks=None, skip_prediction=False) "
                "as Query3DUnified wires it — an arbitrary callable has no backward here and there is no autograd fallback")
        mh_kw = mask_head.keywords
    enc.last_memory_keep = []          # memory-dropout keep masks of this forward, in (layer, group) order (for tests)
    query, query_masks, query_pos = input_dict["query"]
    names = [m for g in enc._program() for m in g]
    masks, flat, nfeat, is_list = [], [], [], []
    for m in names:
        feat, mask, pos = input_dict[m]
        masks.append(mask)
        fl = list(feat[:enc.num_layers]) if isinstance(feat, (list, tuple)) else [feat]
        nfeat.append(len(fl))
        is_list.append(isinstance(feat, (list, tuple)))
        flat += fl + [pos]
    mh_feats = [] if mh is None else [f[0] for f in list(mh_kw["seg_fts_for_match"])[:len(mh.mask_pred_list)]]
    params = list(enc.named_parameters())
    mh_params = [] if mh is None else list(mh.named_parameters())
    first = {id(p): n for n, p in reversed(params)}
    canon = {n: first[id(p)] for n, p in enc.named_parameters(remove_duplicate=False)}
    meta = dict(canon=canon, names=names, masks=masks, query_masks=query_masks, pairwise_locs=pairwise_locs,
                param_names=params, mh=mh, mh_kw=mh_kw, n_mh_feats=len(mh_feats), mh_param_names=mh_params, nfeat=nfeat,
                is_list=is_list)
    outs = DecoderFunction.apply(enc, meta, query, query_pos, *flat, *mh_feats, *[p for _, p in params],
                                 *[p for _, p in mh_params])
    if enc.use_self_mask and meta.get("final_mask") is not None:
        rep = meta["final_mask"].repeat_interleave(enc.num_heads, 0)      # (B*H, N, S), row b*H + h (:84)
        for m in input_dict.keys():                # the reference leaves the last attention mask in input_dict (:85-88)
            if m not in ("query", "prompt"):
                input_dict[m][1] = rep
    if "voxel" in input_dict and isinstance(input_dict["voxel"][0], (list, tuple)):
        input_dict["voxel"][0] = input_dict["voxel"][0][enc.num_layers - 1]      # what the reference's loop leaves (:90-91)
    return outs[0], list(outs[1::2]), list(outs[2::2])
