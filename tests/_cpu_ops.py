"""TEST INFRASTRUCTURE — a torch-on-CPU emulation of the kernels behind `pq3d_b200.ops`, written from the ABI
documentation in include/pq3d_b200.h (pointer + stride semantics included), so that the HOST logic of the product —
operand layouts, grouped-launch arguments, gradient routing, stream-free ordering of `train_engine`, packed-weight
bookkeeping — can be exercised in the `-m "not gpu"` suite, where no kernel can run.  It is never imported by the
package; the product path still raises without the CUDA extension.  bf16 rounding is applied where the kernels round
(operands and bf16 outputs), accumulation is fp32.
"""
from __future__ import annotations

import contextlib
import math

import torch

from pq3d_b200 import ops, query_encoder, rng

bf16 = torch.bfloat16
LN2 = math.log(2.0)


def _as(t, size, stride, extra_offset=0):
    """Strided window starting at t's first element — what a kernel sees from (pointer, strides)."""
    return t.as_strided(size, stride, t.storage_offset() + extra_offset)


def linear(A, W, out, *, M, N, K, bias=None, bias_along_m=False, bias_group_stride=0, groups=1, a_group_rows=0,
           w_group_rows=0, ldc=None, c_group_stride=0, row_zero=None, row_zero_group_stride=0, alpha=1.0, alpha_ncols=0,
           relu=False, block_n=0, max_ctas=0, a_row_offsets=None, w_const=False, no_pairs=False):
    assert A.dtype == bf16 and W.dtype == bf16 and K % 64 == 0
    # the ABI's documented constraints (TMA: 16-byte aligned bases, 16-byte granular pitches >= K)
    assert A.stride(0) % 8 == 0 and W.stride(0) % 8 == 0 and A.stride(0) >= K and W.stride(0) >= K
    assert A.storage_offset() % 8 == 0 and W.storage_offset() % 8 == 0 and A.stride(1) == 1 and W.stride(1) == 1
    ldc = out.stride(-2) if ldc is None else ldc
    lda, ldw = A.stride(0), W.stride(0)
    # TMA zero-fills rows past the operand's extent
    def rows(T, ld, group_rows, n, offsets=None):
        full = torch.zeros(groups, n, K)
        for g in range(groups):
            r0 = g * group_rows if offsets is None else int(offsets[g])
            avail = max(0, min(n, T.shape[0] - r0))
            if avail:
                full[g, :avail] = _as(T, (avail, K), (ld, 1), r0 * ld).float()
        return full
    # TMA stores clip at the tensor map's extent in whole 16-byte chunks: when N * elem_size is not a multiple of 16 the
    # columns up to the next 16-byte boundary are written too (with the epilogue of whatever W rows lie there — zero
    # rows past the operand's extent).  Callers keep pad columns there (the self-attention V^T pitch); emulated so that
    # the teacher-forced parity test sees the same bytes.  (C with a non-16-byte-granular pitch takes direct stores.)
    esz = out.element_size()
    tma_path = (out.storage_offset() * esz) % 16 == 0 and (ldc * esz) % 16 == 0 and (c_group_stride * esz) % 16 == 0
    N_logical = N
    if tma_path and (N * esz) % 16 != 0:
        N = min((N * esz + 15) // 16 * 16 // esz, ldc)
    C = torch.einsum("gmk,gnk->gmn", rows(A, lda, a_group_rows, M, a_row_offsets), rows(W, ldw, w_group_rows, N))
    if bias is not None:
        if bias_along_m:
            b = _as(bias, (groups, M), (bias_group_stride, 1))
            C = C + b[:, :, None]
        else:
            b = torch.zeros(groups, N)
            b[:, :N_logical] = _as(bias, (groups, N_logical), (bias_group_stride, 1))
            C = C + b[:, None, :]
    if alpha_ncols:
        C[:, :, :alpha_ncols] *= alpha
    if relu:
        C = C.relu()
    if row_zero is not None:
        rz = _as(row_zero.view(torch.uint8), (groups, M), (row_zero_group_stride, 1)).bool()
        C = C.masked_fill(rz[:, :, None], 0.0)
    _as(out, (groups, M, N), (c_group_stride, ldc, 1)).copy_(C.to(out.dtype))
    ops._count()
    return out


def bgemm(A, W, C, *, alpha=1.0, block_n=0):
    C.copy_((alpha * torch.einsum("abmk,abnk->abmn", A.float(), W.float())).to(C.dtype))
    ops._count()
    return C


def _unpack_mask(bits, b_stride, h_stride, q_stride, B, H, Nq, S):
    """bool (B, H, Nq, S), True = ignore, from the packed words at word address b*bs + h*hs + n*qs + s/32."""
    W = (S + 31) // 32
    words = _as(bits, (B, H, Nq, W), (b_stride, h_stride, q_stride, 1)).to(torch.int64) & 0xFFFFFFFF
    s = torch.arange(S)
    return ((words[..., s // 32] >> (s % 32)) & 1).bool()


def _scores(Q, ldq, q_col0, K, ldk, k_col0, S, S_pitch, B, H, Nq):
    q = _as(Q, (B, Nq, H, 64), (Nq * ldq, ldq, 64, 1), q_col0).float()
    k = _as(K, (B, S, H, 64), (S_pitch * ldk, ldk, 64, 1), k_col0).float()
    return q, k, torch.einsum("bnhd,bshd->bhns", q, k)


def attention(Q, q_mem_stride, mems, O, o_mem_stride, B, H, Nq, zero_attn, score_bias=None, stats=None, drop_p=0.0,
              seed=None, sites=None):
    ldq, ldo = Q.stride(0), O.stride(-2)
    for i, m in enumerate(mems):
        S = m.S
        q, k, s2 = _scores(Q, ldq, i * q_mem_stride, m.K, m.K.stride(0), m.k_col0, S, m.S_pitch, B, H, Nq)
        ldvt = m.Vt.stride(0)
        v = _as(m.Vt, (H, 64, B, S), (64 * ldvt, ldvt, m.Vt_pitch, 1), m.vt_row0 * ldvt).float()      # (H, d, B, S)
        if score_bias is not None:
            s2 = s2 + score_bias[..., :S]
        if m.mask_bits is not None:
            s2 = s2.masked_fill(_unpack_mask(m.mask_bits, m.mask_b_stride, m.mask_h_stride, m.mask_q_stride, B, H, Nq, S),
                                float("-inf"))
        mx = s2.max(-1).values
        if zero_attn:
            mx = mx.clamp_min(0.0)
        mx = torch.where(torch.isinf(mx), torch.zeros_like(mx), mx)
        # the kernel's softmax REFERENCE per CTA = (scene, head, 128-query tile): memories of more than two 128-key
        # tiles with add_zero_attn take the ONE_PASS schedule, p = ex2(s - 0) (the zero-attn key pins score 0 into every
        # row), unless a row sum leaves the safe range (then the CTA redoes the sweep with the row maxima).  The
        # reference point decides where the bf16 rounding of P falls, so it is part of the rounding-matched emulation.
        T = (S + 127) // 128
        tiles = torch.full((B,), T, dtype=torch.int64)
        if m.kv_tiles is not None:
            tiles = m.kv_tiles.to(torch.int64).clamp(min=1, max=T)
        one_pass = (tiles > 2) & bool(zero_attn)                                     # (B,)
        if bool(one_pass.any()):
            l0 = torch.exp2(s2).sum(-1)                                              # (B, H, Nq)
            qt = (Nq + 127) // 128
            bad = ~(l0 < 1.2676506e30)
            bad = torch.nn.functional.pad(bad, (0, qt * 128 - Nq)).view(B, H, qt, 128).any(-1, keepdim=True)
            bad = bad.expand(B, H, qt, 128).reshape(B, H, qt * 128)[..., :Nq]
            use0 = one_pass.view(B, 1, 1) & ~bad
            mx = torch.where(use0, torch.zeros_like(mx), mx)
        p = torch.exp2(s2 - mx[..., None])
        l = p.sum(-1) + (torch.exp2(-mx) if zero_attn else 0.0)
        if drop_p > 0.0:
            s_pad = (S + 127) // 128 * 128
            e = torch.arange(B * H * Nq).view(B, H, Nq, 1) * s_pad + torch.arange(S).view(1, 1, 1, S)
            keep = rng.keep_mask(int(seed.item()) & 0xFFFFFFFF, sites[i], e, drop_p)
            p = torch.where(keep, p * (1.0 / (1.0 - drop_p)), torch.zeros_like(p))
        o = torch.einsum("bhns,hdbs->bnhd", p.to(bf16).float(), v) / l.permute(0, 2, 1)[..., None].clamp_min(1e-38)
        o = torch.where((l == 0).permute(0, 2, 1)[..., None], torch.zeros_like(o), o)
        _as(O, (B, Nq, H, 64), (Nq * ldo, ldo, 64, 1), i * o_mem_stride).copy_(o.to(bf16))
        if stats is not None:
            stats[0][i].copy_(mx)
            stats[1][i].copy_(l)
    ops._count()
    return O


def attn_delta(dO, O, delta, B, H, N):
    a = _as(dO, (B, N, H, 64), (N * dO.stride(0), dO.stride(0), 64, 1)).float()
    o = _as(O, (B, N, H, 64), (N * dO.stride(0), dO.stride(0), 64, 1)).float()
    delta.copy_((a * o).sum(-1).permute(0, 2, 1))
    ops._count()


def attention_bwd(Q, q_col0, dO, do_col0, K, k_col0, V, v_col0, S, S_pitch, stat_m, stat_l, delta, dK, dk_col0, dV, dv_col0,
                  dQ32, dq_col0, B, H, Nq, *, mask_bits=None, mask_strides=(0, 0, 0), bias=None, dS_out=None, drop_p=0.0,
                  seed=None, site=0):
    assert Nq <= 128
    for t in (Q, dO, K, V, dK, dV):
        assert t.stride(0) % 8 == 0 and t.storage_offset() % 8 == 0 and t.stride(1) == 1
    assert all(c % 8 == 0 for c in (q_col0, do_col0, k_col0, v_col0, dk_col0, dv_col0)) and dq_col0 % 4 == 0
    assert dQ32.stride(0) % 4 == 0 and S_pitch >= S
    q, k, s2 = _scores(Q, Q.stride(0), q_col0, K, K.stride(0), k_col0, S, S_pitch, B, H, Nq)
    v = _as(V, (B, S, H, 64), (S_pitch * V.stride(0), V.stride(0), 64, 1), v_col0).float()
    do = _as(dO, (B, Nq, H, 64), (Nq * dO.stride(0), dO.stride(0), 64, 1), do_col0).float()
    if bias is not None:
        s2 = s2 + bias[..., :S]
    m, l = stat_m.view(B, H, Nq, 1), stat_l.view(B, H, Nq, 1)
    P = torch.where(l > 0, torch.exp2(s2 - m - torch.log2(l.clamp_min(1e-38))), torch.zeros_like(s2))
    if mask_bits is not None:
        P = P.masked_fill(_unpack_mask(mask_bits, *mask_strides, B, H, Nq, S), 0.0)
    ks = torch.ones_like(P)
    if drop_p > 0.0:
        s_pad = (S + 127) // 128 * 128
        e = torch.arange(B * H * Nq).view(B, H, Nq, 1) * s_pad + torch.arange(S).view(1, 1, 1, S)
        ks = rng.keep_mask(int(seed.item()) & 0xFFFFFFFF, site, e, drop_p).float() / (1.0 - drop_p)
    dP = torch.einsum("bnhd,bshd->bhns", do, v)
    dS = (P * (LN2 * dP * ks - LN2 * delta.view(B, H, Nq, 1))).to(bf16).float()
    Pd = (P * ks).to(bf16).float()
    _as(dV, (B, S, H, 64), (S_pitch * dV.stride(0), dV.stride(0), 64, 1), dv_col0).copy_(
        torch.einsum("bhns,bnhd->bshd", Pd, do).to(bf16))
    _as(dK, (B, S, H, 64), (S_pitch * dK.stride(0), dK.stride(0), 64, 1), dk_col0).copy_(
        torch.einsum("bhns,bnhd->bshd", dS, q).to(bf16))
    _as(dQ32, (B, Nq, H, 64), (Nq * dQ32.stride(0), dQ32.stride(0), 64, 1), dq_col0).add_(
        ops.Q_SCALE * torch.einsum("bhns,bshd->bnhd", dS, k))
    if dS_out is not None:
        n_cols = min(dS_out.shape[3], S)
        dS_out[..., :n_cols].copy_(dS[..., :n_cols].to(bf16))
        dS_out[..., n_cols:].zero_()
    ops._count()


def spatial_bias(pairwise_locs, loc_w, loc_b, out):
    L, B, H, N, ld = out.shape
    v = torch.einsum("bnmk,lhk->lbhnm", pairwise_locs, loc_w) + loc_b[:, None, :, None, None]
    out[..., :N].copy_(torch.log2(v.clamp_min(0.0).clamp_min(1e-6)))
    ops._count()


def spatial_bias_bwd(pairwise_locs, loc_w, loc_b, dS, ld, d_w, d_b, B, H, N):
    v = torch.einsum("bnmk,hk->bhnm", pairwise_locs, loc_w) + loc_b[None, :, None, None]
    g = torch.where(v > 1e-6, dS[..., :N].float() / (v * LN2), torch.zeros_like(v))
    d_w.add_(torch.einsum("bhnm,bnmk->hk", g, pairwise_locs))
    d_b.add_(g.sum((0, 2, 3)))
    ops._count()


def ingest_memory(feat, pos, xk, xv, S_pitch):
    B, S, D = feat.shape
    for dst, src in ((xv, feat), (xk, feat if pos is None else feat + pos)):
        if dst is not None:
            d3 = _as(dst, (B, S_pitch, D), (S_pitch * D, D, 1))
            d3.zero_()
            d3[:, :S].copy_(src.to(bf16))
    ops._count()


def ingest_memories(feats, pos, xk, xv, mem_stride, S_pitch):
    B, S, D = feats[0].shape
    for m, f in enumerate(feats):
        for dst, src in ((xv, f), (xk, f + pos)):
            d3 = _as(dst, (B, S_pitch, D), (S_pitch * D, D, 1), m * mem_stride)
            d3.zero_()
            d3[:, :S].copy_(src.to(bf16))
    ops._count()


def _drop_keep(seed, site, G, R, D, p):
    e = torch.arange(G * R * D).view(G, R, D)
    return rng.keep_mask(int(seed.item()) & 0xFFFFFFFF, site, e, p)


def _ln_fwd(y, y_group_stride, residual, gamma, beta, G, eps, R, D, drop_p, seed, site, row_w, rows_per_scene):
    yv = _as(y, (G, R, D), (y_group_stride, D, 1)) if y is not None else torch.zeros(G, R, D)
    if drop_p > 0.0:
        yv = torch.where(_drop_keep(seed, site, G, R, D, drop_p), yv * (1.0 / (1.0 - drop_p)), torch.zeros_like(yv))
    x = yv + (residual.view(1, R, D) if residual is not None else 0.0)
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((x - mean) ** 2).mean(-1, keepdim=True) + eps)
    xhat = (x - mean) * rstd
    if row_w is not None:
        w = row_w.view(-1, G).t().repeat_interleave(rows_per_scene, dim=1).view(G, R, 1)
    else:
        w = torch.full((G, R, 1), 1.0 / G)
    return x, xhat, rstd, w, yv


def _ln_emit(out, pos, out_f32, out_bf16, out_pos_bf16):
    if out_f32 is not None:
        out_f32.view(out.shape).copy_(out)
    if out_bf16 is not None:
        out_bf16.view(out.shape).copy_(out.to(bf16))
    if out_pos_bf16 is not None:
        out_pos_bf16.view(out.shape).copy_((out + pos.view(out.shape)).to(bf16))


def add_layernorm(y, residual, gamma, beta, eps, R, D, G=1, y_group_stride=0, pos=None, out_f32=None, out_bf16=None,
                  out_pos_bf16=None):
    add_layernorm_train(y, residual, gamma, beta, eps, R, D, G, y_group_stride, pos, out_f32, out_bf16, out_pos_bf16)


def add_layernorm_train(y, residual, gamma, beta, eps, R, D, G=1, y_group_stride=0, pos=None, out_f32=None, out_bf16=None,
                        out_pos_bf16=None, drop_p=0.0, seed=None, site=0, row_w=None, rows_per_scene=0):
    _, xhat, _, w, _ = _ln_fwd(y, y_group_stride, residual, gamma, beta, G, eps, R, D, drop_p, seed, site, row_w,
                               rows_per_scene)
    out = (w * (xhat * gamma.view(G, 1, D) + beta.view(G, 1, D))).sum(0)
    _ln_emit(out, pos, out_f32, out_bf16, out_pos_bf16)
    ops._count()


def layernorm_bwd(y, residual, gamma, d_out, eps, R, D, G=1, y_group_stride=0, d_x=None, dx_group_stride=0, d_res=None,
                  d_gamma=None, d_beta=None, d_x16=None, drop_p=0.0, seed=None, site=0, row_w=None, rows_per_scene=0):
    _, xhat, rstd, w, _ = _ln_fwd(y, y_group_stride, residual, gamma, None, G, eps, R, D, drop_p, seed, site, row_w,
                                  rows_per_scene)
    go = d_out.view(1, R, D) * w
    dxh = go * gamma.view(G, 1, D)
    dx = rstd * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))
    dy = dx
    if drop_p > 0.0:
        dy = torch.where(_drop_keep(seed, site, G, R, D, drop_p), dx * (1.0 / (1.0 - drop_p)), torch.zeros_like(dx))
    if d_x is not None:
        _as(d_x, (G, R, D), (dx_group_stride, D, 1)).copy_(dy)
    if d_x16 is not None:
        _as(d_x16, (G, R, D), (dx_group_stride, D, 1)).copy_(dy.to(bf16))
    if d_res is not None:
        d_res.view(R, D).copy_(dx.sum(0))
    if d_gamma is not None:
        d_gamma.view(G, D).add_((go * xhat).sum(1))
    if d_beta is not None:
        d_beta.view(G, D).add_(go.sum(1))
    ops._count()


def pack_mask(mask, bits=None, unmask_full_rows=False, mask_fixed=None, active_tiles=None):
    S = mask.shape[-1]
    W = ops.mask_words(S)
    m = mask.reshape(-1, S).clone()
    if unmask_full_rows:
        m[m.all(-1)] = False
    if mask_fixed is not None:
        mask_fixed.view(torch.bool).view(-1, S).copy_(m)
    padded = torch.ones(m.shape[0], W * 32, dtype=torch.bool)
    padded[:, :S] = m
    words = (padded.view(-1, W, 32).to(torch.int64) << torch.arange(32)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    if bits is None:
        bits = torch.empty(mask.shape[:-1] + (W,), dtype=torch.int32)
    bits.view(-1, W).copy_(words)
    if active_tiles is not None:
        vis = (~padded).view(mask.shape[0], -1, W // 4, 128).any(-1).any(1)            # (batch, tile)
        idx = torch.arange(1, W // 4 + 1)
        active_tiles.copy_((vis * idx).max(-1).values.to(torch.int32))
    ops._count()
    return bits


def cast_bf16(x, out, add=None):
    assert x.numel() % 4 == 0 and x.is_contiguous() and out.is_contiguous()
    out.copy_((x if add is None else x + add).to(bf16))
    ops._count()


def transpose_cast(x, out_t, out_c=None, gate=None, scale=1.0):
    if x.ndim == 2:
        x = x[None, None]
        out_t = None if out_t is None else out_t[None, None]
        out_c = None if out_c is None else out_c[None, None]
        gate = None if gate is None else gate[None, None]
    v = x.float() * scale
    if gate is not None:
        v = torch.where(gate.float() > 0, v, torch.zeros_like(v))
    if out_c is not None:
        out_c.copy_(v.to(bf16))
    if out_t is not None:
        out_t.zero_()
        out_t[..., :x.shape[2]].copy_(v.transpose(-1, -2).to(bf16))
    ops._count()


def colsum(x, out, accumulate=False, gate=None, scale=1.0):
    assert x.shape[1] % 2 == 0 and x.stride(0) % 2 == 0 and x.stride(1) == 1, "pq3d_colsum: C, ld must be even"
    assert (x.storage_offset() * x.element_size()) % 8 == 0
    v = x.float()
    if gate is not None:
        v = torch.where(gate.float() > 0, v, torch.zeros_like(v))
    s = v.sum(0) * scale
    out.copy_(out + s if accumulate else s)
    ops._count()


def add3(a, b, c, out):
    assert out.numel() % 4 == 0 and a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
    out.copy_(a + b + (0 if c is None else c))
    ops._count()


def dropout_bf16(x, drop_p, seed, site):
    keep = rng.keep_mask(int(seed.item()) & 0xFFFFFFFF, site, torch.arange(x.numel()).view(x.shape), drop_p)
    x.copy_(torch.where(keep, x.float() * (1.0 / (1.0 - drop_p)), torch.zeros_like(x, dtype=torch.float32)).to(bf16))
    ops._count()


def gate_mix(gate_logits, query, update, out):
    g = torch.sigmoid(gate_logits)
    out.copy_((1.0 - g) * query + g * update)
    ops._count()


def gate_mix_bwd(gate_logits, query, update, d_out, d_gl, d_gl16, d_update, d_query):
    g = torch.sigmoid(gate_logits)
    dg = d_out * (update - query) * g * (1.0 - g)
    d_gl.copy_(dg)
    d_gl16.copy_(dg.to(bf16))
    d_update.copy_(d_out * g)
    d_query.copy_(d_out * (1.0 - g))
    ops._count()


def mask_head_finalize(raw, mem_mask_ptrs, n_mem, seg_masks, mask_logits, attn_mask, B, S, N, masks=None):
    cnt = (~masks[:n_mem]).sum(0).float()
    v = raw.view(B, S, N) / (cnt[..., None] + 1e-8)
    v = v.masked_fill(seg_masks.view(B, S, 1), -1e6)
    mask_logits.view(B, S, N).copy_(v)
    attn_mask.view(B, N, S).copy_((1.0 / (1.0 + torch.exp(-v)) < 0.5).permute(0, 2, 1))
    ops._count()


def mask_head_finalize_bwd(d_logits, masks, n_mem, d_raw16, B, S, N):
    cnt = (~masks[:n_mem]).sum(0).float()
    v = d_logits.view(B, S, N) / (cnt[..., None] + 1e-8)
    v = v.masked_fill(masks[n_mem].view(B, S, 1), 0.0)
    d_raw16.zero_()
    d_raw16.view(B, S, -1)[..., :N].copy_(v.to(bf16))
    ops._count()


def fourier_pos(xyz, coord_min, coord_max, gauss_B, out):
    B, L = xyz.shape[:2]
    x = (xyz[..., :3].float() - coord_min[:, None, :]) / (coord_max[:, None, :] - coord_min[:, None, :]) * (2 * math.pi)
    proj = (x.reshape(-1, 3) @ gauss_B).view(B * L, -1)
    out.copy_(torch.cat([proj.sin(), proj.cos()], dim=1).to(bf16))
    ops._count()


def pairwise_locs(centers, out=None, eps=1e-10):
    from pq3d_b200 import synth
    r = synth.pairwise_locs_cpu(centers[..., :3].float())
    if out is not None:
        out.copy_(r)
        r = out
    ops._count()
    return r


def _refresh(self):
    """_Packed.refresh without the device-side segment table: the same copies, on the host."""
    for src, dst_c, dst_t, r0, is_f32 in self._segs:
        R, C = src.shape
        if is_f32:
            dst_c.copy_(src)
        else:
            dst_c[r0:r0 + R].copy_(src.to(bf16))
            if dst_t is not None:
                dst_t[:, r0:r0 + R].copy_(src.t().to(bf16))
    self.stale = False
    ops._count()


PATCHED = ["linear", "bgemm", "attention", "attn_delta", "attention_bwd", "spatial_bias", "spatial_bias_bwd", "ingest_memory",
           "ingest_memories",
           "add_layernorm", "add_layernorm_train", "layernorm_bwd", "pack_mask", "cast_bf16", "transpose_cast", "colsum",
           "add3", "dropout_bf16", "gate_mix", "mask_head_finalize", "mask_head_finalize_bwd",
           "fourier_pos", "pairwise_locs", "gate_mix_bwd"]


@contextlib.contextmanager
def cpu_backend():
    """Route `pq3d_b200.ops` through the emulation for the duration of the block."""
    saved = {n: getattr(ops, n) for n in PATCHED}
    saved_refresh, saved_req = query_encoder._Packed.refresh, ops.require_device_tensor
    try:
        for n in PATCHED:
            setattr(ops, n, globals()[n])
        query_encoder._Packed.refresh = _refresh
        ops.require_device_tensor = lambda t: True
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        query_encoder._Packed.refresh = saved_refresh
        ops.require_device_tensor = saved_req
