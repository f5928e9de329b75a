"""Tensor-level wrappers over the C ABI (include/pq3d_b200.h): dtype / contiguity / device checks
live here (the mirror of the reference extension's CHECK_* macros,
modules/third_party/pointnet2/_ext_src/include/utils.h:5-25); the kernels see raw pointers.
All launches go to torch's current CUDA stream, so they are CUDA-graph capturable."""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

bf16 = torch.bfloat16
LOG2E = 1.4426950408889634
Q_SCALE = LOG2E / 8.0     # attention kernels take Q pre-scaled by log2(e)/sqrt(head_dim=64): scores in the log2 domain
LAUNCHES = 0          # kernels launched through this module (bench.py reports it as gpu_launches)


_PRIORITY = [0]
_SHARED_SM = [False]


@contextlib.contextmanager
def shared_sm(on: bool):
    """GEMMs launched inside use the narrow-tile variant that keeps 96 KB (not 192 KB) of operands in flight, so two CTAs
    share an SM: for callers that run several batches concurrently on different streams (pq3d_linear_bf16_ex flags bit 2)."""
    prev = _SHARED_SM[0]
    _SHARED_SM[0] = bool(on)
    try:
        yield
    finally:
        _SHARED_SM[0] = prev


@contextlib.contextmanager
def launch_priority(p: int):
    """Kernels launched inside run with CUDA launch priority p (0 = default = least urgent, negative = more urgent;
    recorded per node when a graph captures them).  pq3d_set_launch_priority."""
    prev = _PRIORITY[0]
    if p != prev:
        _lib.lib().pq3d_set_launch_priority(int(p))
        _PRIORITY[0] = int(p)
    try:
        yield
    finally:
        if _PRIORITY[0] != prev:
            _lib.lib().pq3d_set_launch_priority(prev)
            _PRIORITY[0] = prev


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_device_tensor(t: torch.Tensor) -> bool:
    """True when `t` lives where the kernels can read it (the tests' host-logic emulation replaces this)."""
    return t.is_cuda


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str, ndim: Optional[int] = None):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (pq3d_b200 has no CPU path)")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream; a tensor on another GPU would be touched from the wrong context
        raise ValueError(f"{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()} "
                         "(call torch.cuda.set_device(local_rank) as DDP launchers do)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.ndim != ndim:
        raise ValueError(f"{name} must be {ndim}-D, got shape {tuple(t.shape)}")


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def mask_words(S: int) -> int:
    return (S + 127) // 128 * 4


def linear(A: torch.Tensor, W: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
           bias: Optional[torch.Tensor] = None, bias_along_m: bool = False, bias_group_stride: int = 0,
           groups: int = 1, a_group_rows: int = 0, w_group_rows: int = 0, ldc: Optional[int] = None,
           c_group_stride: int = 0, row_zero: Optional[torch.Tensor] = None, row_zero_group_stride: int = 0,
           alpha: float = 1.0, alpha_ncols: int = 0, relu: bool = False, block_n: int = 0, max_ctas: int = 0,
           a_row_offsets: Optional[Sequence[int]] = None, w_const: bool = False, no_pairs: bool = False) -> torch.Tensor:
    """out[g] = epilogue(A[g] @ W[g].T); A, W are 2-D bf16 views with unit inner stride.  max_ctas / a_row_offsets: see
    pq3d_linear_bf16_ex; w_const=True: W are weights (never written inside the forward), their fetch may start early;
    no_pairs=True: no cta_group::2 clusters (several graphs in flight on different streams)."""
    _chk(A, bf16, "A", 2)
    _chk(W, bf16, "W", 2)
    if A.stride(1) != 1 or W.stride(1) != 1:
        raise ValueError("A and W need unit stride along K")
    if out.dtype not in (bf16, torch.float32):
        raise TypeError("out must be bf16 or fp32")
    if bias is not None:
        _chk(bias, torch.float32, "bias")
    if row_zero is not None and row_zero.dtype not in (torch.bool, torch.uint8):
        raise TypeError("row_zero must be bool/uint8")
    ldc = out.stride(-2) if ldc is None else ldc
    offs = None if a_row_offsets is None else (C.c_int32 * groups)(*[int(v) for v in a_row_offsets])
    rc = _lib.lib().pq3d_linear_bf16_ex(
        A.data_ptr(), A.stride(0), A.shape[0], a_group_rows, W.data_ptr(), W.stride(0), W.shape[0], w_group_rows,
        out.data_ptr(), ldc, c_group_stride, int(out.dtype == torch.float32), _p(bias), bias_group_stride,
        int(bias_along_m), _p(row_zero), row_zero_group_stride, M, N, K, groups, float(alpha), alpha_ncols,
        int(relu), block_n, int(max_ctas), offs, int(w_const) | (2 if no_pairs else 0) | (4 if _SHARED_SM[0] else 0),
        _stream())
    _lib.check(rc, "pq3d_linear_bf16")
    _count()
    return out


def bgemm(A: torch.Tensor, W: torch.Tensor, C: torch.Tensor, *, alpha: float = 1.0, block_n: int = 0) -> torch.Tensor:
    """C[g1, g2] = alpha * A[g1, g2] @ W[g1, g2].T over 4-D strided views (G1, G2, rows, K) with unit inner stride."""
    _chk(A, bf16, "A", 4)
    _chk(W, bf16, "W", 4)
    if C.dtype not in (bf16, torch.float32) or C.ndim != 4:
        raise TypeError("C must be a 4-D bf16 or fp32 tensor")
    G1, G2, M, K = A.shape
    N = W.shape[2]
    if tuple(W.shape) != (G1, G2, N, K) or tuple(C.shape) != (G1, G2, M, N):
        raise ValueError(f"bgemm shape mismatch: A {tuple(A.shape)} W {tuple(W.shape)} C {tuple(C.shape)}")
    if A.stride(3) != 1 or W.stride(3) != 1 or C.stride(3) != 1:
        raise ValueError("bgemm operands need unit inner stride")
    rc = _lib.lib().pq3d_bgemm_bf16(
        A.data_ptr(), A.stride(2), A.stride(1), A.stride(0), W.data_ptr(), W.stride(2), W.stride(1), W.stride(0),
        C.data_ptr(), C.stride(2), C.stride(1), C.stride(0), int(C.dtype == torch.float32), M, N, K, G2, G1,
        float(alpha), block_n, _stream())
    _lib.check(rc, "pq3d_bgemm_bf16")
    _count()
    return C


class AttnMemory:
    """One memory's projected operands for `attention` (see pq3d_attention_fwd)."""
    __slots__ = ("K", "k_col0", "Vt", "vt_row0", "S", "S_pitch", "Vt_pitch", "mask_bits", "mask_b_stride",
                 "mask_h_stride", "mask_q_stride", "kv_tiles")

    def __init__(self, K, k_col0, Vt, vt_row0, S, S_pitch, mask_bits=None, mask_b_stride=0, mask_h_stride=0,
                 mask_q_stride=0, Vt_pitch=None, kv_tiles=None):
        self.kv_tiles = kv_tiles
        self.K, self.k_col0, self.Vt, self.vt_row0 = K, k_col0, Vt, vt_row0
        self.S, self.S_pitch = S, S_pitch
        self.Vt_pitch = S_pitch if Vt_pitch is None else Vt_pitch
        self.mask_bits, self.mask_b_stride = mask_bits, mask_b_stride
        self.mask_h_stride, self.mask_q_stride = mask_h_stride, mask_q_stride


def attention(Q: torch.Tensor, q_mem_stride: int, mems: Sequence[AttnMemory], O: torch.Tensor, o_mem_stride: int,
              B: int, H: int, Nq: int, zero_attn: bool, score_bias: Optional[torch.Tensor] = None,
              stats: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, drop_p: float = 0.0,
              seed: Optional[torch.Tensor] = None, sites: Optional[Sequence[int]] = None) -> torch.Tensor:
    """score_bias: optional fp32 (B, H, Nq, ld) added to the scores (ld = keys padded to a multiple of 128).
    drop_p > 0 (training): dropout on the probabilities, seed = device int32 tensor, sites = one stream id per memory."""
    _chk(Q, bf16, "Q", 2)
    _chk(O, bf16, "O")
    n = len(mems)
    vp = C.c_void_p * n
    i64 = C.c_int64 * n
    i32 = C.c_int32 * n
    for m in mems:
        _chk(m.K, bf16, "K", 2)
        _chk(m.Vt, bf16, "Vt", 2)
        if m.mask_bits is not None:
            _chk(m.mask_bits, torch.int32, "mask_bits")
    bias_ld = 0
    if score_bias is not None:
        _chk(score_bias, torch.float32, "score_bias", 4)
        if not score_bias.is_contiguous() or tuple(score_bias.shape[:3]) != (B, H, Nq):
            raise ValueError("score_bias must be contiguous (B, H, Nq, ld)")
        bias_ld = score_bias.shape[3]
    has_mask = any(m.mask_bits is not None for m in mems)
    train_args = ()
    fn = _lib.lib().pq3d_attention_fwd
    if drop_p > 0.0:
        _chk(seed, torch.int32, "seed")
        fn = _lib.lib().pq3d_attention_fwd_train
        train_args = (float(drop_p), seed.data_ptr(), (C.c_uint32 * n)(*sites))
    rc = fn(
        n, Q.data_ptr(), Q.stride(0), q_mem_stride,
        vp(*[m.K.data_ptr() for m in mems]), i64(*[m.K.stride(0) for m in mems]), i64(*[m.k_col0 for m in mems]),
        vp(*[m.Vt.data_ptr() for m in mems]), i64(*[m.Vt.stride(0) for m in mems]), i64(*[m.vt_row0 for m in mems]),
        i64(*[m.Vt.shape[0] for m in mems]), i32(*[m.S for m in mems]), i32(*[m.S_pitch for m in mems]),
        i32(*[m.Vt_pitch for m in mems]),
        vp(*[_p(m.mask_bits) for m in mems]) if has_mask else None,
        i64(*[m.mask_b_stride for m in mems]), i64(*[m.mask_h_stride for m in mems]),
        i64(*[m.mask_q_stride for m in mems]),
        vp(*[_p(m.kv_tiles) for m in mems]) if any(m.kv_tiles is not None for m in mems) else None,
        O.data_ptr(), O.stride(-2), o_mem_stride, B, H, Nq, int(zero_attn), _p(score_bias), bias_ld,
        None if stats is None else stats[0].data_ptr(), None if stats is None else stats[1].data_ptr(),
        0 if stats is None else B * H * Nq, *train_args, _stream())
    _lib.check(rc, "pq3d_attention_fwd")
    _count()
    return O


def bias_ld(S: int) -> int:
    return (S + 127) // 128 * 128


def spatial_bias(pairwise_locs: torch.Tensor, loc_w: torch.Tensor, loc_b: torch.Tensor, out: torch.Tensor):
    """pairwise_locs (B,N,N,5), loc_w (L,H,5), loc_b (L,H) -> out (L,B,H,N,ld) fp32."""
    _chk(pairwise_locs, torch.float32, "pairwise_locs", 4)
    _chk(loc_w, torch.float32, "loc_w", 3)
    _chk(loc_b, torch.float32, "loc_b", 2)
    _chk(out, torch.float32, "out", 5)
    L, B, H, N, ld = out.shape
    if not (pairwise_locs.is_contiguous() and loc_w.is_contiguous() and loc_b.is_contiguous() and out.is_contiguous()):
        raise ValueError("spatial_bias operands must be contiguous")
    if tuple(pairwise_locs.shape) != (B, N, N, 5) or tuple(loc_w.shape) != (L, H, 5) or tuple(loc_b.shape) != (L, H):
        raise ValueError("spatial_bias shape mismatch")
    rc = _lib.lib().pq3d_spatial_bias(pairwise_locs.data_ptr(), loc_w.data_ptr(), loc_b.data_ptr(), out.data_ptr(),
                                      L, B, H, N, ld, _stream())
    _lib.check(rc, "pq3d_spatial_bias")
    _count()


def ingest_memory(feat: torch.Tensor, pos: Optional[torch.Tensor], xk: Optional[torch.Tensor],
                  xv: Optional[torch.Tensor], S_pitch: int):
    _chk(feat, torch.float32, "feat", 3)
    if not feat.is_contiguous():
        raise ValueError("feat must be contiguous (B, S, D)")
    B, S, D = feat.shape
    if pos is not None:
        _chk(pos, torch.float32, "pos", 3)
        if not pos.is_contiguous() or pos.shape != feat.shape:
            raise ValueError("pos must be contiguous and shaped like feat")
    rc = _lib.lib().pq3d_ingest_memory(feat.data_ptr(), _p(pos), _p(xk), _p(xv), B, S, S_pitch, D, _stream())
    _lib.check(rc, "pq3d_ingest_memory")
    _count()


def ingest_memories(feats: Sequence[torch.Tensor], pos: torch.Tensor, xk: torch.Tensor, xv: torch.Tensor, mem_stride: int,
                    S_pitch: int):
    """Several (B, S, D) fp32 feature tables sharing one positional table -> bf16 operands at xk/xv + m*mem_stride."""
    B, S, D = feats[0].shape
    for f in feats:
        _chk(f, torch.float32, "feat", 3)
        if not f.is_contiguous() or tuple(f.shape) != (B, S, D):
            raise ValueError("feature tables must be contiguous and equally shaped")
    _chk(pos, torch.float32, "pos", 3)
    if not pos.is_contiguous() or pos.shape != feats[0].shape:
        raise ValueError("pos must be contiguous and shaped like the feature tables")
    ptrs = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
    rc = _lib.lib().pq3d_ingest_memories(len(feats), ptrs, pos.data_ptr(), xk.data_ptr(), xv.data_ptr(), mem_stride, B, S,
                                         S_pitch, D, _stream())
    _lib.check(rc, "pq3d_ingest_memories")
    _count()


def add_layernorm(y: Optional[torch.Tensor], residual: Optional[torch.Tensor], gamma: torch.Tensor,
                  beta: torch.Tensor, eps: float, R: int, D: int, G: int = 1, y_group_stride: int = 0,
                  pos: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
                  out_bf16: Optional[torch.Tensor] = None, out_pos_bf16: Optional[torch.Tensor] = None):
    for t, nm in ((y, "y"), (residual, "residual"), (gamma, "gamma"), (beta, "beta"), (pos, "pos"), (out_f32, "out_f32")):
        if t is not None:
            _chk(t, torch.float32, nm)
    for t, nm in ((out_bf16, "out_bf16"), (out_pos_bf16, "out_pos_bf16")):
        if t is not None:
            _chk(t, bf16, nm)
    rc = _lib.lib().pq3d_add_layernorm(_p(y), y_group_stride, _p(residual), gamma.data_ptr(), beta.data_ptr(), G,
                                       float(eps), R, D, _p(pos), _p(out_f32), _p(out_bf16), _p(out_pos_bf16),
                                       _stream())
    _lib.check(rc, "pq3d_add_layernorm")
    _count()


def pack_mask(mask: torch.Tensor, bits: Optional[torch.Tensor] = None, unmask_full_rows: bool = False,
              mask_fixed: Optional[torch.Tensor] = None, active_tiles: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bool (..., S) -> int32 (..., W) packed bits, 1 = ignore.  active_tiles: optional int32 (mask.shape[0],) that
    receives per leading-dim entry the number of 128-key tiles up to the last visible key."""
    if mask.dtype != torch.bool:
        raise TypeError(f"mask must be torch.bool (PyTorch mask convention), got {mask.dtype}")
    if not mask.is_cuda or not mask.is_contiguous():
        raise ValueError("mask must be a contiguous CUDA tensor")
    S = mask.shape[-1]
    rows = mask.numel() // S
    if bits is None:
        bits = torch.empty(mask.shape[:-1] + (mask_words(S),), dtype=torch.int32, device=mask.device)
    rpb = 0
    if active_tiles is not None:
        _chk(active_tiles, torch.int32, "active_tiles", 1)
        if active_tiles.shape[0] != mask.shape[0]:
            raise ValueError("active_tiles must have one entry per leading-dimension index of the mask")
        rpb = rows // mask.shape[0]
    rc = _lib.lib().pq3d_pack_mask(mask.data_ptr(), bits.data_ptr(), rows, S, int(unmask_full_rows),
                                   _p(mask_fixed), _p(active_tiles), rpb, _stream())
    _lib.check(rc, "pq3d_pack_mask")
    _count()
    return bits


def mask_head_finalize(raw: torch.Tensor, mem_mask_ptrs: torch.Tensor, n_mem: int, seg_masks: torch.Tensor,
                       mask_logits: torch.Tensor, attn_mask: torch.Tensor, B: int, S: int, N: int, masks=None):
    """masks: the [n_mem + 1, B, S] bool tensor the pointer table refers to (unused here; the CPU emulation of the
    tests reads it instead of dereferencing device pointers)."""
    _chk(raw, torch.float32, "raw")
    _chk(mask_logits, torch.float32, "mask_logits")
    rc = _lib.lib().pq3d_mask_head_finalize(raw.data_ptr(), mem_mask_ptrs.data_ptr(), n_mem, seg_masks.data_ptr(),
                                            mask_logits.data_ptr(), attn_mask.data_ptr(), B, S, N, _stream())
    _lib.check(rc, "pq3d_mask_head_finalize")
    _count()


def gate_mix(gate_logits, query, update, out):
    rc = _lib.lib().pq3d_gate_mix(gate_logits.data_ptr(), query.data_ptr(), update.data_ptr(), out.data_ptr(),
                                  out.numel(), _stream())
    _lib.check(rc, "pq3d_gate_mix")
    _count()


def cast_bf16(x: torch.Tensor, out: torch.Tensor, add: Optional[torch.Tensor] = None):
    _chk(x, torch.float32, "x")
    _chk(out, bf16, "out")
    rc = _lib.lib().pq3d_cast_bf16(x.data_ptr(), _p(add), out.data_ptr(), x.numel(), _stream())
    _lib.check(rc, "pq3d_cast_bf16")
    _count()


def fourier_pos(xyz: torch.Tensor, coord_min: torch.Tensor, coord_max: torch.Tensor, gauss_B: torch.Tensor,
                out: torch.Tensor):
    """xyz (B, L, >=3) fp32 (last-dim stride 1) -> out (B*L, d_pos) bf16."""
    _chk(xyz, torch.float32, "xyz", 3)
    _chk(out, bf16, "out")
    B, L = xyz.shape[:2]
    if xyz.stride(2) != 1 or xyz.stride(0) != L * xyz.stride(1):
        xyz = xyz.contiguous()
    rc = _lib.lib().pq3d_fourier_pos(xyz.data_ptr(), xyz.stride(1), coord_min.contiguous().data_ptr(),
                                     coord_max.contiguous().data_ptr(), gauss_B.contiguous().data_ptr(),
                                     out.data_ptr(), B, L, 2 * gauss_B.shape[1], _stream())
    _lib.check(rc, "pq3d_fourier_pos")
    _count()


def pairwise_locs(centers: torch.Tensor, out: Optional[torch.Tensor] = None, eps: float = 1e-10) -> torch.Tensor:
    """centers (B, N, >=3) fp32 -> (B, N, N, 5) fp32."""
    _chk(centers, torch.float32, "centers", 3)
    B, N = centers.shape[:2]
    if centers.stride(2) != 1 or centers.stride(0) != N * centers.stride(1):
        centers = centers.contiguous()
    if out is None:
        out = torch.empty(B, N, N, 5, dtype=torch.float32, device=centers.device)
    rc = _lib.lib().pq3d_pairwise_locs(centers.data_ptr(), centers.stride(1), out.data_ptr(), B, N, eps, _stream())
    _lib.check(rc, "pq3d_pairwise_locs")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# backward companions
# ------------------------------------------------------------------------------------------------
def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def transpose_cast(x: torch.Tensor, out_t: Optional[torch.Tensor], out_c: Optional[torch.Tensor] = None,
                   gate: Optional[torch.Tensor] = None, scale: float = 1.0):
    """x: (B1, B2, R, C) strided view (fp32 or bf16, unit inner stride).  out_t: (B1, B2, C, Rp) bf16 receives the
    transpose (columns R..Rp zero), out_c: (B1, B2, R, C) bf16 the cast copy; gate (bf16, like x): ReLU backward."""
    if x.ndim == 2:
        x = x[None, None]
        out_t = None if out_t is None else out_t[None, None]
        out_c = None if out_c is None else out_c[None, None]
        gate = None if gate is None else gate[None, None]
    if x.dtype not in (torch.float32, bf16) or x.stride(3) != 1:
        raise TypeError("transpose_cast input must be fp32/bf16 with unit inner stride")
    B1, B2, R, Cc = x.shape
    Rp = R if out_t is None else out_t.shape[3]
    for t, nm in ((out_t, "out_t"), (out_c, "out_c"), (gate, "gate")):
        if t is not None:
            _chk(t, bf16, nm, 4)
            if t.stride(3) != 1:
                raise ValueError(f"{nm} needs unit inner stride")

    def st(t):
        return (0, 0, 0) if t is None else (t.stride(2), t.stride(0), t.stride(1))
    rc = _lib.lib().pq3d_transpose_cast(x.data_ptr(), int(x.dtype == torch.float32), x.stride(2), x.stride(0), x.stride(1),
                                        _p(gate), *st(gate), _p(out_t), *st(out_t), _p(out_c), *st(out_c), R, Cc, Rp,
                                        B1, B2, float(scale), _stream())
    _lib.check(rc, "pq3d_transpose_cast")
    _count()


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = False, gate: Optional[torch.Tensor] = None,
           scale: float = 1.0):
    """out[c] (+)= sum_r x[r, c] * (gate[r, c] > 0); x fp32 or bf16 2-D view with unit inner stride."""
    if x.dtype not in (torch.float32, bf16) or x.ndim != 2 or x.stride(1) != 1:
        raise TypeError("colsum input must be a 2-D fp32/bf16 view with unit inner stride")
    _chk(out, torch.float32, "out")
    rc = _lib.lib().pq3d_colsum(x.data_ptr(), int(x.dtype == torch.float32), x.stride(0), _p(gate),
                                0 if gate is None else gate.stride(0), out.data_ptr(), x.shape[0], x.shape[1],
                                int(accumulate), float(scale), _stream())
    _lib.check(rc, "pq3d_colsum")
    _count()


def layernorm_bwd(y, residual, gamma, d_out, eps, R, D, G=1, y_group_stride=0, d_x=None, dx_group_stride=0, d_res=None,
                  d_gamma=None, d_beta=None, d_x16=None, drop_p=0.0, seed=None, site=0, row_w=None, rows_per_scene=0):
    if d_x16 is not None:
        _chk(d_x16, bf16, "d_x16")
    if row_w is not None:
        _chk(row_w, torch.float32, "row_w")
    rc = _lib.lib().pq3d_layernorm_bwd(_p(y), y_group_stride, _p(residual), gamma.data_ptr(), d_out.data_ptr(), G,
                                       float(eps), R, D, _p(d_x), dx_group_stride, _p(d_x16), _p(d_res), _p(d_gamma),
                                       _p(d_beta), float(drop_p), _p(seed), site, _p(row_w), rows_per_scene, _stream())
    _lib.check(rc, "pq3d_layernorm_bwd")
    _count()


def attn_delta(dO: torch.Tensor, O: torch.Tensor, delta: torch.Tensor, B: int, H: int, N: int):
    _chk(dO, bf16, "dO", 2)
    _chk(O, bf16, "O", 2)
    rc = _lib.lib().pq3d_attn_delta(dO.data_ptr(), O.data_ptr(), dO.stride(0), delta.data_ptr(), B, H, N, _stream())
    _lib.check(rc, "pq3d_attn_delta")
    _count()


def softmax_bwd(S2, dP, delta, m, l, P, dS, Pt, dSt, B, H, N, S, ld, Np, bias=None, mask_bits=None,
                mask_strides=(0, 0, 0)):
    rc = _lib.lib().pq3d_softmax_bwd(S2.data_ptr(), dP.data_ptr(), delta.data_ptr(), m.data_ptr(), l.data_ptr(),
                                     _p(bias), 0 if bias is None else bias.shape[-1], _p(mask_bits), *mask_strides,
                                     _p(P), dS.data_ptr(), Pt.data_ptr(), dSt.data_ptr(), B, H, N, S, ld, Np,
                                     _stream())
    _lib.check(rc, "pq3d_softmax_bwd")
    _count()


def spatial_bias_bwd(pairwise_locs, loc_w, loc_b, dS, ld, d_w, d_b, B, H, N):
    rc = _lib.lib().pq3d_spatial_bias_bwd(pairwise_locs.data_ptr(), loc_w.data_ptr(), loc_b.data_ptr(), dS.data_ptr(), ld,
                                          d_w.data_ptr(), d_b.data_ptr(), B, H, N, _stream())
    _lib.check(rc, "pq3d_spatial_bias_bwd")
    _count()


def add3(a, b, c, out):
    rc = _lib.lib().pq3d_add3(a.data_ptr(), b.data_ptr(), _p(c), out.data_ptr(), out.numel(), _stream())
    _lib.check(rc, "pq3d_add3")
    _count()


def pack_segments(segs: torch.Tensor, tile_start: torch.Tensor, total_tiles: int):
    """segs: int64 (n_seg, 8) device table, tile_start: int32 (n_seg,) device (see pq3d_pack_segments)."""
    _chk(segs, torch.int64, "segs", 2)
    _chk(tile_start, torch.int32, "tile_start", 1)
    rc = _lib.lib().pq3d_pack_segments(segs.data_ptr(), tile_start.data_ptr(), segs.shape[0], total_tiles, _stream())
    _lib.check(rc, "pq3d_pack_segments")
    _count()


def attention_bwd(Q, q_col0, dO, do_col0, K, k_col0, V, v_col0, S, S_pitch, stat_m, stat_l, delta, dK, dk_col0, dV,
                  dv_col0, dQ32, dq_col0, B, H, Nq, *, mask_bits=None, mask_strides=(0, 0, 0), bias=None, dS_out=None,
                  drop_p=0.0, seed=None, site=0):
    """Fused attention backward (pq3d_attention_bwd).  Q, dO, K, V, dK, dV: 2-D bf16 with unit inner stride; dQ32: 2-D
    fp32, accumulated (zero-fill first); stats / delta: fp32 (B, H, Nq)."""
    for t, nm in ((Q, "Q"), (dO, "dO"), (K, "K"), (V, "V"), (dK, "dK"), (dV, "dV")):
        _chk(t, bf16, nm, 2)
        if t.stride(1) != 1:
            raise ValueError(f"{nm} needs unit inner stride")
    _chk(dQ32, torch.float32, "dQ32", 2)
    for t, nm in ((stat_m, "stat_m"), (stat_l, "stat_l"), (delta, "delta")):
        _chk(t, torch.float32, nm)
        if not t.is_contiguous() or t.numel() != B * H * Nq:
            raise ValueError(f"{nm} must be contiguous (B, H, Nq)")
    bias_ld = 0
    if bias is not None:
        _chk(bias, torch.float32, "bias", 4)
        bias_ld = bias.shape[3]
    ds_ld = 0
    if dS_out is not None:
        _chk(dS_out, bf16, "dS_out", 4)
        ds_ld = dS_out.shape[3]
    rc = _lib.lib().pq3d_attention_bwd(
        Q.data_ptr(), Q.stride(0), q_col0, dO.data_ptr(), dO.stride(0), do_col0, K.data_ptr(), K.stride(0), k_col0,
        V.data_ptr(), V.stride(0), v_col0, S, S_pitch, _p(mask_bits), *mask_strides, _p(bias), bias_ld,
        stat_m.data_ptr(), stat_l.data_ptr(), delta.data_ptr(), dK.data_ptr(), dK.stride(0), dk_col0, dV.data_ptr(),
        dV.stride(0), dv_col0, dQ32.data_ptr(), dQ32.stride(0), dq_col0, _p(dS_out), ds_ld, B, H, Nq, float(Q_SCALE),
        float(drop_p), _p(seed), site, _stream())
    _lib.check(rc, "pq3d_attention_bwd")
    _count()


def add_layernorm_train(y, residual, gamma, beta, eps, R, D, G=1, y_group_stride=0, pos=None, out_f32=None, out_bf16=None,
                        out_pos_bf16=None, drop_p=0.0, seed=None, site=0, row_w=None, rows_per_scene=0):
    """Training-mode add_layernorm: dropout on y (counter RNG, device seed tensor) and per-scene memory weights."""
    for t, nm in ((y, "y"), (residual, "residual"), (gamma, "gamma"), (beta, "beta"), (pos, "pos"), (out_f32, "out_f32"),
                  (row_w, "row_w")):
        if t is not None:
            _chk(t, torch.float32, nm)
    for t, nm in ((out_bf16, "out_bf16"), (out_pos_bf16, "out_pos_bf16")):
        if t is not None:
            _chk(t, bf16, nm)
    if seed is not None:
        _chk(seed, torch.int32, "seed")
    rc = _lib.lib().pq3d_add_layernorm_train(_p(y), y_group_stride, _p(residual), gamma.data_ptr(), beta.data_ptr(), G,
                                             float(eps), R, D, _p(pos), _p(out_f32), _p(out_bf16), _p(out_pos_bf16),
                                             float(drop_p), _p(seed), site, _p(row_w), rows_per_scene, _stream())
    _lib.check(rc, "pq3d_add_layernorm_train")
    _count()


def dropout_bf16(x: torch.Tensor, drop_p: float, seed: torch.Tensor, site: int):
    _chk(x, bf16, "x")
    _chk(seed, torch.int32, "seed")
    if not x.is_contiguous():
        raise ValueError("dropout_bf16 works in place on a contiguous tensor")
    rc = _lib.lib().pq3d_dropout_bf16(x.data_ptr(), x.numel(), float(drop_p), seed.data_ptr(), site, _stream())
    _lib.check(rc, "pq3d_dropout_bf16")
    _count()


def mask_head_finalize_bwd(d_logits: torch.Tensor, masks: torch.Tensor, n_mem: int, d_raw16: torch.Tensor, B: int, S: int,
                           N: int):
    """d_logits fp32 (B, S, N) contiguous, masks bool [n_mem + 1, B, S] contiguous -> d_raw16 bf16 [B*S, Np]."""
    _chk(d_logits, torch.float32, "d_logits")
    _chk(d_raw16, bf16, "d_raw16", 2)
    if not d_logits.is_contiguous() or not masks.is_contiguous() or not d_raw16.is_contiguous():
        raise ValueError("mask_head_finalize_bwd operands must be contiguous")
    rc = _lib.lib().pq3d_mask_head_finalize_bwd(d_logits.data_ptr(), masks.data_ptr(), n_mem, d_raw16.data_ptr(), B, S, N,
                                                d_raw16.shape[1], _stream())
    _lib.check(rc, "pq3d_mask_head_finalize_bwd")
    _count()


def gate_mix_bwd(gate_logits, query, update, d_out, d_gl, d_gl16, d_update, d_query):
    for t, nm in ((gate_logits, "gate_logits"), (query, "query"), (update, "update"), (d_out, "d_out"), (d_gl, "d_gl"),
                  (d_update, "d_update"), (d_query, "d_query")):
        _chk(t, torch.float32, nm)
        if not t.is_contiguous():
            raise ValueError(f"{nm} must be contiguous")
    _chk(d_gl16, bf16, "d_gl16")
    rc = _lib.lib().pq3d_gate_mix_bwd(gate_logits.data_ptr(), query.data_ptr(), update.data_ptr(), d_out.data_ptr(),
                                      d_gl.data_ptr(), d_gl16.data_ptr(), d_update.data_ptr(), d_query.data_ptr(),
                                      d_out.numel(), _stream())
    _lib.check(rc, "pq3d_gate_mix_bwd")
    _count()


# ------------------------------------------------------------------------------------------------
# §8f-2: voxel -> segment pooling
# ------------------------------------------------------------------------------------------------
def segment_csr(p2s: torch.Tensor, voxel_offsets: Sequence[int], max_seg: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """p2s: int64 (Nv_total,) per-scene segment id of every voxel (scenes concatenated); voxel_offsets: B+1 host prefix
    sums.  Returns (perm int32 (Nv_total,), offsets int32 (B*max_seg + 1,)): voxels grouped by (scene, segment), ascending
    inside a segment."""
    _chk(p2s, torch.int64, "point2segment", 1)
    if not p2s.is_contiguous():
        raise ValueError("point2segment must be contiguous")
    B = len(voxel_offsets) - 1
    if voxel_offsets[0] != 0 or voxel_offsets[-1] != p2s.numel():
        raise ValueError("voxel_offsets must run from 0 to the number of voxels")
    offs = (C.c_int64 * (B + 1))(*[int(v) for v in voxel_offsets])
    need = _lib.lib().pq3d_segment_csr_workspace_bytes(offs, B, int(max_seg))
    if need < 0:
        raise ValueError(f"segment_csr: unsupported shape (B={B}, max_seg={max_seg})")
    ws = torch.empty(max(int(need), 4), dtype=torch.uint8, device=p2s.device)
    perm = torch.empty(max(p2s.numel(), 1), dtype=torch.int32, device=p2s.device)
    offsets = torch.empty(B * max_seg + 1, dtype=torch.int32, device=p2s.device)
    rc = _lib.lib().pq3d_segment_csr(p2s.data_ptr(), offs, B, int(max_seg), perm.data_ptr(), offsets.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "pq3d_segment_csr")
    _count(4)
    return perm, offsets


def segment_mean(feat: torch.Tensor, perm: torch.Tensor, offsets: torch.Tensor, out32: Optional[torch.Tensor] = None,
                 out16: Optional[torch.Tensor] = None):
    """feat fp32 (Nv_total, C) -> mean per (scene, segment): out32 fp32 (G, C) and / or out16 bf16 (G, K16 >= C, zero padded)."""
    _chk(feat, torch.float32, "feat", 2)
    _chk(perm, torch.int32, "perm", 1)
    _chk(offsets, torch.int32, "offsets", 1)
    if feat.stride(1) != 1:
        raise ValueError("feat needs unit inner stride")
    G, Cc = offsets.numel() - 1, feat.shape[1]
    if out32 is not None:
        _chk(out32, torch.float32, "out32", 2)
    if out16 is not None:
        _chk(out16, bf16, "out16", 2)
    rc = _lib.lib().pq3d_segment_mean(feat.data_ptr(), feat.stride(0), perm.data_ptr(), offsets.data_ptr(), G, Cc,
                                      _p(out32), 0 if out32 is None else out32.stride(0), _p(out16),
                                      0 if out16 is None else out16.stride(0), 0 if out16 is None else out16.shape[1],
                                      _stream())
    _lib.check(rc, "pq3d_segment_mean")
    _count()


# ------------------------------------------------------------------------------------------------
# §8f-3: matcher cost matrices, matched mask losses
# ------------------------------------------------------------------------------------------------
def match_cost(pred_masks: torch.Tensor, pred_logits: torch.Tensor, tgt_masks: torch.Tensor, tgt_labels: torch.Tensor,
               tgt_count: torch.Tensor, w_class: float, w_mask: float, w_dice: float, ignore_label: int) -> torch.Tensor:
    """pred_masks fp32 (B, S, N), pred_logits fp32 (B, N, C), tgt_masks uint8 (B, Mmax, S), tgt_labels int64 (B, Mmax),
    tgt_count int32 (B,) -> cost fp32 (B, N, Mmax)."""
    _chk(pred_masks, torch.float32, "pred_masks", 3)
    _chk(pred_logits, torch.float32, "pred_logits", 3)
    _chk(tgt_masks, torch.uint8, "tgt_masks", 3)
    _chk(tgt_labels, torch.int64, "tgt_labels", 2)
    _chk(tgt_count, torch.int32, "tgt_count", 1)
    for t in (pred_masks, pred_logits, tgt_masks, tgt_labels):
        if not t.is_contiguous():
            raise ValueError("match_cost operands must be contiguous")
    B, S, N = pred_masks.shape
    Cc, Mmax = pred_logits.shape[2], tgt_masks.shape[1]
    if tuple(pred_logits.shape[:2]) != (B, N) or tuple(tgt_masks.shape) != (B, Mmax, S) or tuple(tgt_labels.shape) != (B, Mmax):
        raise ValueError("match_cost shape mismatch")
    cost = torch.empty(B, N, Mmax, dtype=torch.float32, device=pred_masks.device)
    rc = _lib.lib().pq3d_match_cost(pred_masks.data_ptr(), pred_logits.data_ptr(), tgt_masks.data_ptr(),
                                    tgt_labels.data_ptr(), tgt_count.data_ptr(), cost.data_ptr(), B, N, S, Cc, Mmax,
                                    float(w_class), float(w_mask), float(w_dice), int(ignore_label), _stream())
    _lib.check(rc, "pq3d_match_cost")
    _count()
    return cost


def matched_mask_loss_fwd(pred_masks, tgt_masks, pairs):
    _chk(pred_masks, torch.float32, "pred_masks", 3)
    _chk(tgt_masks, torch.uint8, "tgt_masks", 3)
    _chk(pairs, torch.int32, "pairs", 2)
    B, S, N = pred_masks.shape
    n = pairs.shape[0]
    dev = pred_masks.device
    ce, dice = torch.empty(n, dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    sums = torch.empty(n, 2, dtype=torch.float32, device=dev)
    rc = _lib.lib().pq3d_matched_mask_loss_fwd(pred_masks.data_ptr(), tgt_masks.data_ptr(), pairs.data_ptr(), n, B, N, S,
                                               tgt_masks.shape[1], ce.data_ptr(), dice.data_ptr(), sums.data_ptr(), _stream())
    _lib.check(rc, "pq3d_matched_mask_loss_fwd")
    _count()
    return ce, dice, sums


def matched_mask_loss_bwd(pred_masks, tgt_masks, pairs, g_ce, g_dice, sums):
    B, S, N = pred_masks.shape
    d_pred = torch.zeros_like(pred_masks)
    rc = _lib.lib().pq3d_matched_mask_loss_bwd(pred_masks.data_ptr(), tgt_masks.data_ptr(), pairs.data_ptr(), pairs.shape[0],
                                               B, N, S, tgt_masks.shape[1], g_ce.data_ptr(), g_dice.data_ptr(),
                                               sums.data_ptr(), d_pred.data_ptr(), _stream())
    _lib.check(rc, "pq3d_matched_mask_loss_bwd")
    _count()
    return d_pred
