"""GPU kernel checks, one per process so a trapping kernel cannot poison the others.

    python tests/kernel_checks.py list
    python tests/kernel_checks.py <check-name>        # exit 0 = pass

Each check feeds a C-ABI kernel bf16-rounded random operands and compares against the same math in
plain torch fp32 on the SAME rounded operands, so the only differences are accumulation order and
the output rounding.  Tolerances (||a-b||inf / ||b||inf):
  fp32 outputs : 1e-3 (north-star tolerance; observed ~1e-6)
  bf16 outputs : 2^-8 = 3.9e-3 (one bf16 ulp of the largest element — the output format's own resolution)
  bool / packed-bit outputs : exact
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from pq3d_b200 import ops

DEV = "cuda"
TOL_F32 = 1e-3
TOL_BF16 = 2.0 ** -8


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def rnd(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------------------- GEMM
def _gemm_case(M, N, K, block_n, out_dtype=torch.bfloat16, bias=True, bias_along_m=False, relu=False, alpha=1.0,
               alpha_ncols=0, groups=1, row_zero=False, ldc_pad=0, seed=0):
    g = gen(seed)
    A = rnd((groups * M, K), g).bfloat16()
    W = rnd((groups * N, K), g, 0.05).bfloat16()
    bvec = rnd((groups, M if bias_along_m else N), g) if bias else None
    rz = (torch.rand(groups, M, generator=g) < 0.3).to(DEV) if row_zero else None
    ldc = N + ldc_pad
    out = torch.full((groups, M, ldc), float("nan"), dtype=out_dtype, device=DEV)
    ops.linear(A, W, out, M=M, N=N, K=K, bias=bvec, bias_along_m=bias_along_m,
               bias_group_stride=(bvec.shape[1] if bias and groups > 1 else 0), groups=groups,
               a_group_rows=M if groups > 1 else 0, w_group_rows=N if groups > 1 else 0, ldc=ldc,
               c_group_stride=M * ldc, row_zero=rz, row_zero_group_stride=M if groups > 1 else 0,
               alpha=alpha, alpha_ncols=alpha_ncols, relu=relu, block_n=block_n)
    torch.cuda.synchronize()
    ref = torch.einsum("gmk,gnk->gmn", A.float().view(groups, M, K).double(), W.float().view(groups, N, K).double())
    if bias:
        ref = ref + (bvec.double()[:, :, None] if bias_along_m else bvec.double()[:, None, :])
    if alpha_ncols:
        ref[:, :, :alpha_ncols] *= alpha
    if relu:
        ref = ref.clamp_min(0)
    if row_zero:
        ref = ref.masked_fill(rz[:, :, None], 0.0)
    got = out[:, :, :N]
    assert not torch.isnan(got.float()).any(), "unwritten / NaN outputs"
    if ldc_pad:
        assert torch.isnan(out[:, :, N:].float()).all(), "wrote outside the N columns"
    e = rel(got, ref)
    tol = TOL_F32 if out_dtype == torch.float32 else TOL_BF16
    print(f"gemm M={M} N={N} K={K} bn={block_n} {out_dtype} groups={groups}: rel {e:.2e} (tol {tol:.1e})")
    assert e <= tol


def check_gemm_min():
    _gemm_case(128, 64, 64, 64, torch.float32, bias=False)
    _gemm_case(128, 128, 128, 128, torch.float32)
    _gemm_case(128, 256, 192, 256, torch.float32)


def check_gemm_shapes():
    _gemm_case(400, 768, 768, 64, torch.float32, seed=1)
    _gemm_case(400, 2304, 768, 64, torch.bfloat16, alpha=0.125, alpha_ncols=2304, seed=2)
    _gemm_case(1000, 1536, 768, 128, torch.bfloat16, seed=3)
    _gemm_case(2048, 3072, 768, 256, torch.bfloat16, seed=4)
    _gemm_case(400, 768, 2048, 64, torch.float32, seed=5)          # FFN2: long K, ring wraps
    _gemm_case(400, 2048, 768, 0, torch.bfloat16, relu=True, seed=6)  # auto tile
    _gemm_case(333, 1536, 768, 128, torch.bfloat16, alpha=0.125, alpha_ncols=768, seed=7)


def check_gemm_pair():
    """Problems with more than two waves of 128x256 tiles run on CTA pairs (cta_group::2, 256-row tiles)."""
    _gemm_case(8192, 3072, 768, 256, torch.bfloat16, seed=40)
    _gemm_case(640, 15360, 128, 256, torch.bfloat16, seed=41)                      # odd number of row tiles
    _gemm_case(4224 - 50, 768, 768, 256, torch.bfloat16, groups=3, seed=42)        # grouped, ragged last pair
    _gemm_case(3072, 8192, 768, 256, torch.bfloat16, bias_along_m=True, seed=43)   # V^T form
    _gemm_case(8192, 1024, 256, 256, torch.float32, relu=True, seed=44)


def check_gemm_scheduling_knobs():
    """Launch priority (pq3d_set_launch_priority) and the multi-wave / capped grids (max_ctas < 0 / > 0) change scheduling
    only: results are bit-identical to the default launch."""
    g = gen(77)
    M, N, K = 4096, 1536, 768
    A, W, b = rnd((M, K), g).bfloat16(), rnd((N, K), g, 0.05).bfloat16(), rnd((N,), g)
    outs = []
    for prio, ctas in ((0, 0), (-2, 0), (0, -3), (-1, 40)):
        out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
        with ops.launch_priority(prio):
            ops.linear(A, W, out, M=M, N=N, K=K, bias=b, block_n=256, max_ctas=ctas, no_pairs=True)
        outs.append(out)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert ops._PRIORITY[0] == 0
    print("gemm: launch priority / multi-wave / capped grids bit-identical")
    # the 96 KB-ring variant of the narrow tiles (two CTAs per SM; ops.shared_sm) computes the same numbers
    for bn, (M, N, K) in ((64, (400, 768, 2048)), (128, (1000, 1536, 768)), (64, (400, 2304, 768))):
        A, W, b = rnd((M, K), g).bfloat16(), rnd((N, K), g, 0.05).bfloat16(), rnd((N,), g)
        res = []
        for on in (False, True):
            out = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
            with ops.shared_sm(on):
                ops.linear(A, W, out, M=M, N=N, K=K, bias=b, block_n=bn, w_const=True)
            res.append(out)
        torch.cuda.synchronize()
        assert torch.equal(res[0], res[1]), f"half-ring GEMM differs (bn={bn})"
    print("gemm: half-ring (shared-SM) tiles bit-identical")


def check_bgemm():
    """Strided batched GEMM: per-(scene, head) slices of packed [tokens, heads*64] tensors, ragged M / N per group."""
    g = gen(50)
    B, H, N, S, D = 3, 4, 100, 300, 256
    Q = rnd((B * N, D), g).bfloat16()                 # [B*N, H*64]
    Kt = rnd((B * S, 2 * D), g).bfloat16()            # [B*S, L*H*64], use layer 1
    # scores[b,h] = Q_bh [N,64] @ K_bh[S,64]^T  (fp32 out)
    A = Q.view(B, N, H, 64).permute(0, 2, 1, 3)
    W = Kt.view(B, S, 2, H, 64)[:, :, 1].permute(0, 2, 1, 3)
    Sc = torch.full((B, H, N, 304), float("nan"), device=DEV)
    ops.bgemm(A, W, Sc[..., :S].as_strided((B, H, N, S), Sc.stride()), alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * torch.einsum("bhnd,bhsd->bhns", A.double(), W.double())
    e = rel(Sc[..., :S], ref)
    assert torch.isnan(Sc[..., S:]).all(), "wrote past N"
    print(f"bgemm scores (fp32): rel {e:.2e}")
    assert e <= TOL_F32
    # dV[b,h] = P^T [S, Np] @ dO^T[64, Np]^T with Np = 128 (zero padded contraction), bf16 out into a packed [B*S, D] tensor
    Np = 128
    Pt = torch.zeros(B, H, S, Np, device=DEV, dtype=torch.bfloat16)
    Pt[..., :N] = rnd((B, H, S, N), g, 0.1).bfloat16()
    dOt = torch.zeros(B, H, 64, Np, device=DEV, dtype=torch.bfloat16)
    dOt[..., :N] = rnd((B, H, 64, N), g).bfloat16()
    dV = torch.full((B * S, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.bgemm(Pt, dOt, dV.view(B, S, H, 64).permute(0, 2, 1, 3))
    torch.cuda.synchronize()
    ref = torch.einsum("bhsn,bhdn->bhsd", Pt.double(), dOt.double())
    e = rel(dV.view(B, S, H, 64).permute(0, 2, 1, 3), ref)
    print(f"bgemm dV (bf16, strided store): rel {e:.2e}")
    assert e <= TOL_BF16


def check_gemm_epilogues():
    _gemm_case(3072, 520, 768, 128, torch.bfloat16, bias_along_m=True, seed=8)     # V^T form
    _gemm_case(400, 201, 768, 64, torch.float32, seed=9)                           # cls head: unaligned N
    _gemm_case(300, 100, 2304, 128, torch.float32, groups=4, bias=False, seed=10)  # mask logits per scene
    _gemm_case(400, 768, 768, 64, torch.float32, groups=3, seed=11)                # per-memory out-proj
    _gemm_case(500, 768, 768, 128, torch.bfloat16, bias=False, row_zero=True, seed=12)
    _gemm_case(130, 72, 128, 64, torch.bfloat16, ldc_pad=8, seed=13)


# -------------------------------------------------------------------------------------- attention
def _attn_ref(Q, K, V, mask, zero_attn, bias=None):
    """Q (B,H,N,64) pre-scaled INTO THE LOG2 DOMAIN (kernel contract), K/V (B,H,S,64), mask (B,H,N,S) bool,
    bias in log2 units; fp64 math."""
    s = torch.einsum("bhnd,bhsd->bhns", Q.double(), K.double())
    if bias is not None:
        s = s + bias.double()
    s = s * math.log(2.0)
    s = s.masked_fill(mask, float("-inf"))
    if zero_attn:
        s = torch.cat([s, torch.zeros_like(s[..., :1])], -1)
    p = torch.softmax(s, -1)
    if zero_attn:
        p = p[..., :-1]
    return torch.einsum("bhns,bhsd->bhnd", p, V.double())


def _attn_case(B, H, Nq, S_list, mask_kind, zero_attn=True, spatial=False, seed=0, scale=1.0):
    g = gen(seed)
    n_mem = len(S_list)
    D = H * 64
    Q = rnd((B * Nq, n_mem * D), g, scale).bfloat16()
    O = torch.full((n_mem, B * Nq, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    mems, refs = [], []
    pw = lw = lb = None
    if spatial:
        pw = torch.rand(B, Nq, Nq, 5, generator=g).to(DEV)
        lw = rnd((H, 5), g, 0.7)
        lb = rnd((H,), g, 0.3)
    for i, S in enumerate(S_list):
        Sp = ops.pad8(S)
        L = 2                                    # pretend two layers are stacked; use layer 1
        Kb = rnd((B * Sp, L * D), g, scale).bfloat16()
        Vt = rnd((L * D, B * Sp), g).bfloat16()
        if mask_kind == "kpm":
            m = (torch.rand(B, S, generator=g) < 0.3).to(DEV)
            bits = ops.pack_mask(m)
            strides = (bits.stride(0), 0, 0)
            full = m[:, None, None, :].expand(B, H, Nq, S)
        elif mask_kind == "attn":
            m = (torch.rand(B, Nq, S, generator=g) < 0.5).to(DEV)
            m[:, 3] = True                      # a fully masked row
            bits = ops.pack_mask(m)
            strides = (bits.stride(0), 0, bits.stride(1))
            full = m[:, None].expand(B, H, Nq, S)
        elif mask_kind == "attn_heads":
            m = (torch.rand(B * H, Nq, S, generator=g) < 0.5).to(DEV)
            bits = ops.pack_mask(m)
            strides = (H * bits.stride(0), bits.stride(0), bits.stride(1))
            full = m.view(B, H, Nq, S)
        else:
            bits, strides = None, (0, 0, 0)
            full = torch.zeros(B, H, Nq, S, dtype=torch.bool, device=DEV)
        mems.append(ops.AttnMemory(Kb, D, Vt, D, S, Sp, bits, *strides))
        Qh = Q.view(B, Nq, n_mem, H, 64)[:, :, i].permute(0, 2, 1, 3).float()
        Kh = Kb.view(B, Sp, L, H, 64)[:, :S, 1].permute(0, 2, 1, 3).float()
        Vh = Vt.view(L, H, 64, B, Sp)[1, :, :, :, :S].permute(2, 0, 3, 1).float()
        bias = None
        if spatial:
            loc = torch.relu(torch.einsum("bnmd,hd->bhnm", pw, lw) + lb[None, :, None, None])
            bias = torch.log2(loc.clamp_min(1e-6))
        refs.append(_attn_ref(Qh, Kh, Vh, full, zero_attn, bias))
    sbias = None
    if spatial:
        sbias = torch.full((1, B, H, Nq, ops.bias_ld(Nq)), float("nan"), device=DEV)
        ops.spatial_bias(pw, lw[None].contiguous(), lb[None].contiguous(), sbias)
        sbias = sbias[0]
    ops.attention(Q, D, mems, O, O.stride(0), B, H, Nq, zero_attn, sbias)
    torch.cuda.synchronize()
    for i in range(n_mem):
        got = O[i].view(B, Nq, H, 64).permute(0, 2, 1, 3).float()
        ref = refs[i]
        ok_rows = ~torch.isnan(ref).any(-1)          # (no zero-attn and all masked -> NaN in both)
        assert not torch.isnan(got[ok_rows]).any(), "NaN in attention output"
        # P is rounded to bf16 before the PV product and O is stored as bf16: two bf16 roundings
        e = ((got[ok_rows] - ref[ok_rows]).abs().max() / ref[ok_rows].abs().max()).item()
        print(f"attn B={B} H={H} Nq={Nq} S={S_list[i]} mask={mask_kind} zero={zero_attn} spatial={spatial}: rel {e:.2e}")
        assert e <= 2 * TOL_BF16


def check_attn_basic():
    _attn_case(1, 1, 100, [128], "none", seed=1)
    _attn_case(2, 2, 100, [300], "kpm", seed=2)
    _attn_case(2, 12, 100, [1000], "kpm", seed=3, scale=1.5)


def check_attn_masks():
    _attn_case(2, 3, 100, [260], "attn", seed=4)
    _attn_case(2, 3, 64, [150], "attn_heads", seed=5)
    _attn_case(1, 2, 200, [333], "attn", seed=6)          # two query tiles
    _attn_case(2, 2, 100, [32], "kpm", seed=7)            # prompt-sized memory
    _attn_case(2, 4, 100, [520, 520, 520], "kpm", seed=8)  # three memories in one launch


def check_attn_spatial():
    _attn_case(2, 12, 100, [100], "kpm", zero_attn=False, spatial=True, seed=9)
    _attn_case(1, 2, 37, [37], "none", zero_attn=False, spatial=True, seed=10)
    _attn_case(1, 2, 200, [200], "kpm", zero_attn=False, spatial=True, seed=14)   # two query tiles, two resident key tiles


def check_attn_resident():
    """Memories of <= 256 keys keep their score tiles in TMEM (single QK^T pass)."""
    _attn_case(2, 3, 100, [256], "kpm", seed=15)
    _attn_case(2, 3, 100, [129], "attn", seed=16)
    _attn_case(2, 3, 100, [257], "kpm", seed=17)           # just past the resident limit: streaming path


def check_attn_ragged_tiles():
    """pack_mask's active-tile counts let the attention kernel skip trailing padding tiles; results must not change."""
    g = gen(30)
    B, H, Nq, S = 3, 2, 100, 1000
    D = H * 64
    lens = [1000, 130, 400]
    m = torch.arange(S)[None, :] >= torch.tensor(lens)[:, None]
    m = (m | (torch.rand(B, S, generator=g) < 0.1)).to(DEV)
    tiles = torch.full((B,), -1, dtype=torch.int32, device=DEV)
    bits = ops.pack_mask(m, active_tiles=tiles)
    torch.cuda.synchronize()
    assert tiles.tolist() == [8, 2, 4], tiles.tolist()
    Q = rnd((B * Nq, D), g).bfloat16()
    Sp = ops.pad8(S)
    Kb, Vt = rnd((B * Sp, D), g).bfloat16(), rnd((D, B * Sp), g).bfloat16()
    outs = []
    for kt in (None, tiles):
        O = torch.full((1, B * Nq, D), float("nan"), dtype=torch.bfloat16, device=DEV)
        ops.attention(Q, 0, [ops.AttnMemory(Kb, 0, Vt, 0, S, Sp, bits, bits.stride(0), 0, 0, kv_tiles=kt)], O,
                      B * Nq * D, B, H, Nq, True)
        torch.cuda.synchronize()
        outs.append(O)
    assert not torch.isnan(outs[1].float()).any()
    # a trimmed scene may take a different schedule (<= 2 tiles: resident, running max) than the full-length
    # sweep (one pass, reference 0): same math, different rounding -> compare within the bf16 output resolution
    assert rel(outs[1], outs[0]) <= 2 * TOL_BF16, "skipping fully masked tiles changed the result"
    # per-query masks: tile count is the max over the scene's rows; the all-masked fix-up makes a row fully visible
    am = torch.ones(2, 5, 300, dtype=torch.bool, device=DEV)
    am[0, :, :100] = False
    am[1, 0, :10] = False            # rows 1..4 of scene 1 stay fully masked
    t2 = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.pack_mask(am, active_tiles=t2)
    t3 = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.pack_mask(am, unmask_full_rows=True, active_tiles=t3)
    torch.cuda.synchronize()
    assert t2.tolist() == [1, 1] and t3.tolist() == [1, 3], (t2.tolist(), t3.tolist())
    print(f"ragged tiles: counts exact, outputs equal within {rel(outs[1], outs[0]):.1e} with and without skipping")


def check_attn_long():
    _attn_case(1, 2, 100, [4096], "kpm", seed=11, scale=2.0)


def check_attn_schedules():
    """Long zero-attn memories take the one-pass schedule (reference score 0, no running max); row sums
    beyond 2^100 make the CTA redo the sweep with running maxima; both must match the fp64 softmax."""
    from pq3d_b200 import _lib
    _attn_case(2, 4, 100, [700], "kpm", seed=18, scale=1.0)              # one pass
    _attn_case(2, 4, 100, [700], "attn", seed=19, scale=1.0)
    _attn_case(2, 4, 100, [700], "kpm", seed=20, scale=4.5)              # scores ~ +-160 (log2): overflow -> redo
    _lib.lib().pq3d_debug_force_two_pass(1)
    try:
        _attn_case(2, 4, 100, [700], "kpm", seed=18, scale=1.0)          # forced running-max schedule
    finally:
        _lib.lib().pq3d_debug_force_two_pass(0)
    _attn_case(2, 4, 100, [700], "kpm", zero_attn=False, seed=21)        # no zero-attn: running-max schedule


# ------------------------------------------------------------------------------------ elementwise
def check_ingest():
    g = gen(20)
    for (B, S, D, with_pos) in [(2, 100, 768, True), (3, 37, 768, False), (1, 2048, 768, True)]:
        Sp = ops.pad8(S)
        feat, pos = rnd((B, S, D), g), (rnd((B, S, D), g) if with_pos else None)
        xk = torch.full((B, Sp, D), float("nan"), dtype=torch.bfloat16, device=DEV)
        xv = torch.full_like(xk, float("nan"))
        ops.ingest_memory(feat, pos, xk, xv, Sp)
        torch.cuda.synchronize()
        assert torch.equal(xv[:, :S], feat.bfloat16())
        assert torch.equal(xk[:, :S], (feat + pos).bfloat16() if with_pos else feat.bfloat16())
        assert (xv[:, S:] == 0).all() and (xk[:, S:] == 0).all()
    print("ingest: exact")


def check_add_layernorm():
    g = gen(21)
    for (G, R, D, eps) in [(1, 400, 768, 1e-5), (3, 400, 768, 1e-5), (1, 77, 768, 1e-12), (1, 50, 384, 1e-12)]:
        y, res, pos = rnd((G, R, D), g), rnd((R, D), g), rnd((R, D), g)
        gamma, beta = 1 + 0.1 * rnd((G, D), g), 0.1 * rnd((G, D), g)
        o32 = torch.empty(R, D, device=DEV)
        ob = torch.empty(R, D, dtype=torch.bfloat16, device=DEV)
        op = torch.empty_like(ob)
        ops.add_layernorm(y, res, gamma, beta, eps, R, D, G=G, y_group_stride=R * D, pos=pos, out_f32=o32,
                          out_bf16=ob, out_pos_bf16=op)
        torch.cuda.synchronize()
        ref = sum(torch.nn.functional.layer_norm((res + y[i]).double(), (D,), gamma[i].double(), beta[i].double(), eps)
                  for i in range(G)) / G
        e = rel(o32, ref)
        print(f"add_layernorm G={G} R={R} D={D}: rel {e:.2e}")
        assert e <= 1e-5
        assert torch.equal(ob, o32.bfloat16()) and torch.equal(op, (o32 + pos).bfloat16())
    # plain LayerNorm of one tensor (no residual): the cls-head LN
    y = rnd((1, 10, 768), g)
    o32 = torch.empty(10, 768, device=DEV)
    ones, zeros = torch.ones(1, 768, device=DEV), torch.zeros(1, 768, device=DEV)
    ops.add_layernorm(y, None, ones, zeros, 1e-12, 10, 768, out_f32=o32)
    torch.cuda.synchronize()
    assert rel(o32, torch.nn.functional.layer_norm(y[0].double(), (768,), eps=1e-12)) <= 1e-5


def check_pack_mask():
    g = gen(22)
    for (rows, S) in [((4,), 100), ((2, 50), 333), ((3, 7), 128), ((2, 5), 4096)]:
        m = (torch.rand(*rows, S, generator=g) < 0.5)
        m[..., 0, :] = True                       # fully masked rows
        md = m.to(DEV)
        bits = ops.pack_mask(md)
        fixed = torch.empty_like(md)
        bits_fix = ops.pack_mask(md, unmask_full_rows=True, mask_fixed=fixed.view(torch.uint8))
        torch.cuda.synchronize()
        W = ops.mask_words(S)

        def expect(mm):
            pad = torch.ones(*mm.shape[:-1], W * 32, dtype=torch.bool)
            pad[..., :S] = mm
            w = pad.view(*mm.shape[:-1], W, 32).long()
            return (w << torch.arange(32)).sum(-1)
        got = bits.cpu().long() & 0xFFFFFFFF
        assert torch.equal(got, expect(m)), "pack mismatch"
        m2 = m.clone()
        m2[m2.all(-1)] = False
        assert torch.equal(bits_fix.cpu().long() & 0xFFFFFFFF, expect(m2)), "pack+fixup mismatch"
        assert torch.equal(fixed.cpu(), m2)
    print("pack_mask: exact")


def check_mask_head_finalize():
    g = gen(23)
    B, S, N, n_mem = 2, 150, 40, 3
    raw = rnd((B, S, N), g, 3.0)
    raw[0, 0, :5] = torch.tensor([0.0, -0.0, 1e-9, -1e-9, -1e-6])
    mem_masks = [(torch.rand(B, S, generator=g) < 0.3).to(DEV) for _ in range(n_mem)]
    mem_masks[0][0, 5] = mem_masks[1][0, 5] = mem_masks[2][0, 5] = True          # no memory valid
    seg = (torch.rand(B, S, generator=g) < 0.1).to(DEV)
    ptrs = torch.tensor([m.data_ptr() for m in mem_masks], dtype=torch.int64, device=DEV)
    logits = torch.empty(B, S, N, device=DEV)
    am = torch.empty(B, N, S, dtype=torch.bool, device=DEV)
    ops.mask_head_finalize(raw, ptrs, n_mem, seg, logits, am, B, S, N)
    torch.cuda.synchronize()
    valid = sum(m[..., None].logical_not() for m in mem_masks)
    ref = raw / (valid + 1e-8)                    # modules/heads/mask_head.py:36
    ref[seg] = -1e6
    ref_am = ref.sigmoid().permute(0, 2, 1) < 0.5
    assert torch.equal(logits, ref), f"mask logits differ: {(logits - ref).abs().max()}"
    assert torch.equal(am, ref_am), f"attn mask differs in {(am != ref_am).sum()} places"
    print("mask_head_finalize: bit-exact vs torch")


# ------------------------------------------------------------------------------------ backward pieces
def check_bwd_elementwise():
    g = gen(60)
    # transpose_cast with relu gate, 2-level batches over packed heads
    B, N, H = 3, 100, 4
    x = rnd((B * N, H * 64), g)
    gate = rnd((B * N, H * 64), g).bfloat16()
    xv = x.view(B, N, H, 64).permute(0, 2, 1, 3)
    gv = gate.view(B, N, H, 64).permute(0, 2, 1, 3)
    out_t = torch.full((B, H, 64, 128), float("nan"), dtype=torch.bfloat16, device=DEV)
    out_c = torch.full((B, H, N, 64), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.transpose_cast(xv, out_t, out_c, gate=gv, scale=0.5)
    torch.cuda.synchronize()
    ref = (0.5 * xv * (gv.float() > 0)).bfloat16()
    assert torch.equal(out_c, ref) and torch.equal(out_t[..., :N], ref.transpose(2, 3)) and (out_t[..., N:] == 0).all()
    # 2-D fp32 -> bf16 transpose with padding (wgrad operand)
    y = rnd((400, 768), g)
    yt = torch.full((768, 448), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.transpose_cast(y, yt)
    torch.cuda.synchronize()
    assert torch.equal(yt[:, :400], y.bfloat16().t()) and (yt[:, 400:] == 0).all()
    # colsum
    cs = torch.empty(768, device=DEV)
    ops.colsum(y, cs)
    ops.colsum(y, cs, accumulate=True)
    torch.cuda.synchronize()
    assert rel(cs, 2 * y.double().sum(0)) <= 1e-5
    # add3
    a, b, c = rnd((400, 768), g), rnd((400, 768), g), rnd((400, 768), g)
    o = torch.empty_like(a)
    ops.add3(a, b, c, o)
    torch.cuda.synchronize()
    assert torch.equal(o, a + b + c)
    print("transpose_cast / colsum / add3: exact")


def check_layernorm_bwd():
    g = gen(61)
    for (G, R, D) in [(1, 400, 768), (3, 400, 768), (1, 77, 384)]:
        y = rnd((G, R, D), g).requires_grad_(True)
        res = rnd((R, D), g).requires_grad_(True)
        gamma = (1 + 0.1 * rnd((G, D), g)).requires_grad_(True)
        beta = (0.1 * rnd((G, D), g)).requires_grad_(True)
        d_out = rnd((R, D), g)
        out = sum(torch.nn.functional.layer_norm((res + y[i]).double(), (D,), gamma[i].double(), beta[i].double(), 1e-5)
                  for i in range(G)) / G
        out.backward(d_out.double())
        d_x = torch.empty(G, R, D, device=DEV)
        d_res = torch.empty(R, D, device=DEV)
        d_g, d_b = torch.zeros(G, D, device=DEV), torch.zeros(G, D, device=DEV)
        ops.layernorm_bwd(y.detach(), res.detach(), gamma.detach(), d_out, 1e-5, R, D, G=G, y_group_stride=R * D, d_x=d_x,
                          dx_group_stride=R * D, d_res=d_res, d_gamma=d_g, d_beta=d_b)
        torch.cuda.synchronize()
        e = max(rel(d_x, y.grad), rel(d_res, res.grad), rel(d_g, gamma.grad), rel(d_b, beta.grad))
        print(f"layernorm_bwd G={G} R={R} D={D}: rel {e:.2e}")
        assert e <= 1e-4


def check_attention_bwd():
    """The attention backward as the decoder composes it: forward kernel (saves m, l), score recompute and the four
    gradient products on the batched GEMM, softmax backward — against autograd of the fp64 softmax attention."""
    g = gen(62)
    B, H, N, S = 2, 3, 100, 300
    D, Sp, Np, ld = H * 64, ops.pad8(300), 128, ops.pad64(300)
    ln2 = math.log(2.0)
    Q2 = rnd((B * N, D), g).bfloat16()                       # log2-domain queries
    Kb = rnd((B * Sp, D), g).bfloat16()
    Vb = rnd((B * Sp, D), g).bfloat16()
    kpm = (torch.rand(B, S, generator=g) < 0.2).to(DEV)
    bits = ops.pack_mask(kpm)
    Vt = torch.zeros(D, B * Sp, dtype=torch.bfloat16, device=DEV)
    Vt.copy_(Vb.t())
    Kt = Kb.t().contiguous()
    O = torch.empty(1, B * N, D, dtype=torch.bfloat16, device=DEV)
    st_m, st_l = torch.empty(B, H, N, device=DEV), torch.empty(B, H, N, device=DEV)
    ops.attention(Q2, 0, [ops.AttnMemory(Kb, 0, Vt, 0, S, Sp, bits, bits.stride(0), 0, 0)], O, B * N * D, B, H, N, True,
                  stats=(st_m, st_l))
    dO = rnd((B * N, D), g).bfloat16()
    # reference (fp64 autograd)
    q = Q2.double().view(B, N, H, 64).permute(0, 2, 1, 3).requires_grad_(True)
    k = Kb.double().view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3).requires_grad_(True)
    v = Vb.double().view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3).requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) * ln2
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    s = torch.cat([s, torch.zeros_like(s[..., :1])], -1)
    pr = torch.softmax(s, -1)[..., :-1]
    o_ref = pr @ v
    o_ref.backward(dO.double().view(B, N, H, 64).permute(0, 2, 1, 3))
    # ours
    Qv = Q2.view(B, N, H, 64).permute(0, 2, 1, 3)
    Kv = Kb.view(B, Sp, H, 64).permute(0, 2, 1, 3)[:, :, :S]
    S2 = torch.zeros(B, H, N, ld, device=DEV)
    ops.bgemm(Qv, Kv, S2[..., :S].as_strided((B, H, N, S), S2.stride()))
    dOv = dO.view(B, N, H, 64).permute(0, 2, 1, 3)
    Vv = Vb.view(B, Sp, H, 64).permute(0, 2, 1, 3)[:, :, :S]
    dP = torch.zeros(B, H, N, ld, device=DEV)
    ops.bgemm(dOv, Vv, dP[..., :S].as_strided((B, H, N, S), dP.stride()))
    delta = torch.empty(B, H, N, device=DEV)
    ops.attn_delta(dO, O[0], delta, B, H, N)
    P = torch.empty(B, H, N, ld, dtype=torch.bfloat16, device=DEV)
    dS = torch.empty_like(P)
    Pt = torch.empty(B, H, ld, Np, dtype=torch.bfloat16, device=DEV)
    dSt = torch.empty_like(Pt)
    ops.softmax_bwd(S2, dP, delta, st_m, st_l, P, dS, Pt, dSt, B, H, N, S, ld, Np, mask_bits=bits,
                    mask_strides=(bits.stride(0), 0, 0))
    dOt = torch.empty(B, H, 64, Np, dtype=torch.bfloat16, device=DEV)
    Q2t = torch.empty(B, H, 64, Np, dtype=torch.bfloat16, device=DEV)
    ops.transpose_cast(dOv, dOt)
    ops.transpose_cast(Qv, Q2t)
    dV = torch.zeros(B * Sp, D, dtype=torch.bfloat16, device=DEV)
    dK = torch.zeros(B * Sp, D, dtype=torch.bfloat16, device=DEV)
    dQ = torch.empty(B * N, D, dtype=torch.bfloat16, device=DEV)
    ops.bgemm(Pt[:, :, :S], dOt, dV.view(B, Sp, H, 64).permute(0, 2, 1, 3)[:, :, :S])
    ops.bgemm(dSt[:, :, :S], Q2t, dK.view(B, Sp, H, 64).permute(0, 2, 1, 3)[:, :, :S])
    Ktv = Kt.view(H, 64, B, Sp).permute(2, 0, 1, 3)            # (B, H, 64, Sp): K^T per head
    Ktp = torch.zeros(B, H, 64, ld, dtype=torch.bfloat16, device=DEV)
    Ktp[..., :S] = Ktv[..., :S]
    ops.bgemm(dS, Ktp, dQ.view(B, N, H, 64).permute(0, 2, 1, 3))
    torch.cuda.synchronize()
    e_p = rel(P[..., :S], pr)
    e_v = rel(dV.view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3), v.grad)
    e_k = rel(dK.view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3), k.grad)
    e_q = rel(dQ.view(B, N, H, 64).permute(0, 2, 1, 3), q.grad)
    print(f"attention backward: P {e_p:.2e}  dV {e_v:.2e}  dK {e_k:.2e}  dQ {e_q:.2e}")
    assert max(e_p, e_v, e_k, e_q) <= 4 * TOL_BF16       # P, dS and the outputs are each rounded to bf16


def _attn_bwd_fused_case(B, H, N, S, mask_kind, zero_attn, with_bias, seed):
    g = gen(seed)
    D, Sp = H * 64, ops.pad8(S)
    ln2 = math.log(2.0)
    Q2 = rnd((B * N, D), g).bfloat16()
    Kb = rnd((B * Sp, D), g).bfloat16()
    Vb = rnd((B * Sp, D), g).bfloat16()
    dO = rnd((B * N, D), g).bfloat16()
    if mask_kind == "kpm":
        mask = (torch.rand(B, S, generator=g) < 0.2).to(DEV)
        bits = ops.pack_mask(mask)
        strides = (bits.stride(0), 0, 0)
        full = mask[:, None, None, :].expand(B, H, N, S)
    else:                                   # per-query mask shared by the heads: (B, N, S)
        mask = (torch.rand(B, N, S, generator=g) < 0.3).to(DEV)
        mask[:, 0] = True                   # a fully masked row
        bits = ops.pack_mask(mask)
        strides = (bits.stride(0), 0, bits.stride(1))
        full = mask[:, None].expand(B, H, N, S)
    bias = None
    if with_bias:
        bias = torch.zeros(B, H, N, ops.bias_ld(S), device=DEV)
        bias[..., :S] = rnd((B, H, N, S), g)
    Vt = torch.zeros(D, B * Sp, dtype=torch.bfloat16, device=DEV)
    Vt.copy_(Vb.t())
    O = torch.empty(1, B * N, D, dtype=torch.bfloat16, device=DEV)
    st_m, st_l = torch.empty(B, H, N, device=DEV), torch.empty(B, H, N, device=DEV)
    ops.attention(Q2, 0, [ops.AttnMemory(Kb, 0, Vt, 0, S, Sp, bits, *strides)], O, B * N * D, B, H, N, zero_attn, bias,
                  stats=(st_m, st_l))
    # reference (fp64 autograd)
    q = Q2.double().view(B, N, H, 64).permute(0, 2, 1, 3).requires_grad_(True)
    k = Kb.double().view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3).requires_grad_(True)
    v = Vb.double().view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3).requires_grad_(True)
    s = q @ k.transpose(-1, -2)
    if with_bias:
        s = s + bias[..., :S].double()
    s = (s * ln2).masked_fill(full, float("-inf"))
    if zero_attn:
        s = torch.cat([s, torch.zeros_like(s[..., :1])], -1)
    pr = torch.softmax(s, -1)
    pr = torch.nan_to_num(pr, nan=0.0)[..., :S]
    o_ref = pr @ v
    o_ref.backward(dO.double().view(B, N, H, 64).permute(0, 2, 1, 3))
    # ours
    delta = torch.empty(B, H, N, device=DEV)
    ops.attn_delta(dO, O[0], delta, B, H, N)
    dK = torch.zeros(B * Sp, D, dtype=torch.bfloat16, device=DEV)
    dV = torch.zeros(B * Sp, D, dtype=torch.bfloat16, device=DEV)
    dQ = torch.zeros(B * N, D, device=DEV)
    dS = torch.full((B, H, N, ops.pad64(S)), float("nan"), dtype=torch.bfloat16, device=DEV) if with_bias else None
    ops.attention_bwd(Q2, 0, dO, 0, Kb, 0, Vb, 0, S, Sp, st_m, st_l, delta, dK, 0, dV, 0, dQ, 0, B, H, N, mask_bits=bits,
                      mask_strides=strides, bias=bias, dS_out=dS)
    torch.cuda.synchronize()
    e_v = rel(dV.view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3), v.grad)
    e_k = rel(dK.view(B, Sp, H, 64)[:, :S].permute(0, 2, 1, 3), k.grad)
    e_q = rel(dQ.view(B, N, H, 64).permute(0, 2, 1, 3) / ops.Q_SCALE, q.grad)
    print(f"fused attention backward B={B} H={H} N={N} S={S} {mask_kind} zero_attn={zero_attn} bias={with_bias}: "
          f"dV {e_v:.2e}  dK {e_k:.2e}  dQ {e_q:.2e}")
    assert max(e_v, e_k, e_q) <= 4 * TOL_BF16
    if dS is not None:
        assert torch.isfinite(dS[..., :S].float()).all()


def check_attention_bwd_fused():
    """pq3d_attention_bwd (one kernel, MN-major operand re-use) against fp64 autograd of the softmax attention."""
    _attn_bwd_fused_case(2, 3, 100, 300, "kpm", True, False, 70)        # cross-attention, key padding, 3 key tiles
    _attn_bwd_fused_case(1, 2, 128, 1100, "kpm", True, False, 71)       # 9 tiles -> split over chunks, dQ atomics
    _attn_bwd_fused_case(2, 2, 48, 200, "attn", True, False, 72)        # per-query masks, a fully masked row
    _attn_bwd_fused_case(2, 3, 100, 100, "kpm", False, True, 73)        # self-attention shape: score bias, no zero-attn


CHECKS = {k[6:]: v for k, v in list(globals().items()) if k.startswith("check_")}

if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "list":
        print("\n".join(CHECKS))
        sys.exit(0)
    torch.manual_seed(0)
    CHECKS[sys.argv[1]]()
    print("PASS", sys.argv[1])
