"""ctypes binding of include/pq3d_b200.h.  The product path has no fallback: if the shared library
is missing (or a call fails) this raises, loudly."""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_C", "libpq3d_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "pq3d_b200.h")

_i64, _i32, _f32, _vp = C.c_int64, C.c_int, C.c_float, C.c_void_p
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)
_pi32 = C.POINTER(C.c_int32)

SIGNATURES: Dict[str, list] = {
    "pq3d_linear_bf16": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _i32,
                         _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _vp],
    "pq3d_linear_bf16_ex": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _i32,
                            _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _i32, _i32, _pi32, _i32, _vp],
    "pq3d_bgemm_bf16": [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32,
                        _i32, _i32, _f32, _i32, _vp],
    "pq3d_attention_fwd": [_i32, _vp, _i64, _i64, _pp, _pi64, _pi64, _pp, _pi64, _pi64, _pi64, _pi32, _pi32, _pi32,
                           _pp, _pi64, _pi64, _pi64, _pp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _i64, _vp],
    "pq3d_attention_fwd_train": [_i32, _vp, _i64, _i64, _pp, _pi64, _pi64, _pp, _pi64, _pi64, _pi64, _pi32, _pi32, _pi32,
                                 _pp, _pi64, _pi64, _pi64, _pp, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp,
                                 _i64, _f32, _vp, C.POINTER(C.c_uint32), _vp],
    "pq3d_spatial_bias": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp],
    "pq3d_debug_set_timeline": [_vp],
    "pq3d_set_launch_priority": [C.c_int],
    "pq3d_debug_set_attention_timeline": [_vp],
    "pq3d_debug_force_two_pass": [_i32],
    "pq3d_transpose_cast": [_vp, _i32, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64,
                            _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "pq3d_colsum": [_vp, _i32, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _f32, _vp],
    "pq3d_layernorm_bwd": [_vp, _i64, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _f32, _vp,
                           C.c_uint32, _vp, _i32, _vp],
    "pq3d_add_layernorm_train": [_vp, _i64, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _f32, _vp,
                                 C.c_uint32, _vp, _i32, _vp],
    "pq3d_dropout_bf16": [_vp, _i64, _f32, _vp, C.c_uint32, _vp],
    "pq3d_attn_delta": [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp],
    "pq3d_softmax_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i32,
                         _i32, _i32, _i32, _i32, _vp],
    "pq3d_spatial_bias_bwd": [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp],
    "pq3d_add3": [_vp, _vp, _vp, _vp, _i64, _vp],
    "pq3d_pack_segments": [_vp, _vp, _i32, _i32, _vp],
    "pq3d_gate_mix_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp],
    "pq3d_mask_head_finalize_bwd": [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp],
    "pq3d_attention_bwd": [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _i64,
                           _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64,
                           _i32, _i32, _i32, _f32, _f32, _vp, C.c_uint32, _vp],
    "pq3d_ingest_memory": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "pq3d_ingest_memories": [_i32, _pp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp],
    "pq3d_add_layernorm": [_vp, _i64, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "pq3d_pack_mask": [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _i64, _vp],
    "pq3d_mask_head_finalize": [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "pq3d_gate_mix": [_vp, _vp, _vp, _vp, _i64, _vp],
    "pq3d_cast_bf16": [_vp, _vp, _vp, _i64, _vp],
    "pq3d_fourier_pos": [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "pq3d_pairwise_locs": [_vp, _i32, _vp, _i32, _i32, _f32, _vp],
    "pq3d_segment_csr_workspace_bytes": [_pi64, _i32, _i32],
    "pq3d_segment_csr": [_vp, _pi64, _i32, _i32, _vp, _vp, _vp, _i64, _vp],
    "pq3d_segment_mean": [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp],
    "pq3d_match_cost": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _i64, _vp],
    "pq3d_matched_mask_loss_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "pq3d_matched_mask_loss_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
    "pq3d_class_probs": [_vp, _vp, _i32, _i32, _vp],
    "pq3d_topk": [_vp, _i32, _i32, _vp, _vp, _vp],
    "pq3d_bincount": [_vp, _i64, _vp, _i32, _vp],
    "pq3d_instseg_scores": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "pq3d_split_index": [_vp, _vp, _i32, _i32, _vp, _vp, _vp],
    "pq3d_instseg_fullres": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp],
}
RESTYPES = {"pq3d_segment_csr_workspace_bytes": _i64}


def declared_symbols() -> List[str]:
    """Every function include/pq3d_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pq3d_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"pq3d_b200: CUDA extension {LIB_PATH} is missing — run `python -m pq3d_b200.build` "
                "(there is no CPU or PyTorch fallback for the decoder kernels)")
        l = C.CDLL(LIB_PATH)
        l.pq3d_last_error.restype = C.c_char_p
        l.pq3d_last_error.argtypes = []
        l.pq3d_abi_version.restype = C.c_int
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = RESTYPES.get(name, C.c_int)
        _lib = l
    return _lib


class Pq3dError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        raise Pq3dError(f"{what} failed (code {rc}): {lib().pq3d_last_error().decode()}")
