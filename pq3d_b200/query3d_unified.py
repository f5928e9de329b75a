"""Query3DUnified — the drop-in model boundary (model/query3d_unified.py:29-238).

Same constructor (`Query3DUnified(cfg)`), `forward(data_dict) -> data_dict` (mutates and returns the
same dict) and `get_opt_params()`; same attribute / parameter names so stage-1 / stage-2 checkpoints
load.  `cfg` may be an omegaconf DictConfig, a plain nested dict or any attribute mapping.

In scope here: coordinate encoders, ObjectEncoder projection
branches, prompt features supplied as `data_dict['prompt_feat']`, pairwise geometry, the decoder,
the mask / ground heads.  Out of scope (raise, never silently fall back): MinkowskiEngine voxel
backbone, PointNet++ tokenizer, CLIP / T5 towers.
"""
from __future__ import annotations

import os
import weakref
from copy import copy
from functools import partial
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .mask_head import MaskHeadSegLevel, MlpHeadRunner, mlp_head_params
from .query_encoder import QueryEncoder, QueryMaskEncoder

bf16 = torch.bfloat16


# ---- cfg access that works for DictConfig / dict / attr-dict --------------------------------
def cfg_get(cfg: Any, key: str, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    if hasattr(cfg, "get"):
        try:
            return cfg.get(key, default)
        except TypeError:
            pass
    return getattr(cfg, key, default)


def cfg_path(cfg: Any, path: str, default=None):
    cur = cfg
    for k in path.split("."):
        cur = cfg_get(cur, k, None)
        if cur is None:
            return default
    return cur


def cfg_plain(cfg: Any):
    """cfg2dict (common/type_utils.py:6-7) without requiring omegaconf."""
    try:
        from omegaconf import OmegaConf, DictConfig, ListConfig      # type: ignore
        if isinstance(cfg, (DictConfig, ListConfig)):
            return OmegaConf.to_container(cfg, resolve=True)
    except Exception:
        pass
    if isinstance(cfg, dict) or (hasattr(cfg, "keys") and hasattr(cfg, "__getitem__")):
        return {k: cfg_plain(cfg[k]) for k in cfg.keys()}
    if isinstance(cfg, (list, tuple)):
        return [cfg_plain(v) for v in cfg]
    return cfg


PROMPT_TYPE_TXT, PROMPT_TYPE_LOC = 1, 3          # data/datasets/constant.py:628-631 (PromptType)


class LinearLN(nn.Sequential):
    """nn.Sequential(Linear, LayerNorm) evaluated by the GEMM + LayerNorm kernels."""

    def __init__(self, d_in, d_out):
        super().__init__(nn.Linear(d_in, d_out), nn.LayerNorm(d_out))
        self._key, self._w = None, None

    def _weights(self, dev):
        ps = list(self.parameters())
        key = (str(dev), tuple(p._version for p in ps), tuple(p.data_ptr() for p in ps))
        if self._key != key:
            w = self[0].weight.detach()
            k_pad = (w.shape[1] + 63) // 64 * 64           # tensor-core K granularity; zero columns are exact
            if k_pad != w.shape[1]:
                w = torch.nn.functional.pad(w, (0, k_pad - w.shape[1]))
            self._w = dict(w=w.to(dev, bf16).contiguous(), b=self[0].bias.detach().float().to(dev),
                           g=self[1].weight.detach().float().to(dev)[None].contiguous(),
                           be=self[1].bias.detach().float().to(dev)[None].contiguous(), eps=self[1].eps, k=k_pad)
            self._key = key
        return self._w

    def run16(self, x16: torch.Tensor, R: int, out32: torch.Tensor, emit=None):
        """emit = (xv16, xk16, pos32): the LayerNorm epilogue also writes the decoder's bf16 operands of this memory,
        xv = bf16(out) and xk = bf16(out + pos) (no fp32 round trip through a separate ingest pass)."""
        w = self._weights(x16.device)
        D = w["w"].shape[0]
        y = torch.empty(R, D, dtype=torch.float32, device=x16.device)
        ops.linear(x16, w["w"], y, M=R, N=D, K=w["k"], bias=w["b"])
        if emit is None:
            ops.add_layernorm(y, None, w["g"], w["be"], w["eps"], R, D, out_f32=out32)
        else:
            xv16, xk16, pos32 = emit
            ops.add_layernorm(y, None, w["g"], w["be"], w["eps"], R, D, pos=pos32, out_f32=out32, out_bf16=xv16,
                              out_pos_bf16=xk16)
        return out32

    def _needs_grad(self, x=None) -> bool:
        return torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters())
                                            or (x is not None and x.requires_grad))

    def forward(self, x: torch.Tensor, emit=None) -> torch.Tensor:
        if self._needs_grad(x):            # training: forward + backward composed from the kernels (train_blocks.py)
            from .train_blocks import linear_ln_train
            return linear_ln_train(x, self[0], self[1])
        lead, din = x.shape[:-1], x.shape[-1]
        R = x.numel() // din
        w = self._weights(x.device)
        x2 = x.reshape(R, din).float()
        if w["k"] != din:
            x2 = torch.nn.functional.pad(x2, (0, w["k"] - din))
        x16 = torch.empty(R, w["k"], dtype=bf16, device=x.device)
        ops.cast_bf16(x2.contiguous(), x16)
        out = torch.empty(R, w["w"].shape[0], dtype=torch.float32, device=x.device)
        return self.run16(x16, R, out, emit=emit).view(*lead, -1)


class CoordinateEncoder(nn.Module):
    """model/query3d_unified.py:15-27: Fourier features (fp32, Gaussian matrix buffer `pos_enc.gauss_B`)
    -> Linear -> LayerNorm."""

    class _PosEnc(nn.Module):
        def __init__(self, d_pos, d_in=3, gauss_scale=1.0):
            super().__init__()
            self.register_buffer("gauss_B", torch.empty(d_in, d_pos // 2).normal_() * gauss_scale)

    def __init__(self, hidden_size, use_projection=True):
        super().__init__()
        if not use_projection:
            raise NotImplementedError("CoordinateEncoder is always built with use_projection=True on this path")
        self.pos_enc = CoordinateEncoder._PosEnc(hidden_size)
        self.feat_proj = LinearLN(hidden_size, hidden_size)

    def forward(self, coords, input_range):
        B, L = coords.shape[:2]
        D = self.pos_enc.gauss_B.shape[1] * 2
        x16 = torch.empty(B * L, D, dtype=bf16, device=coords.device)
        ops.fourier_pos(coords.float(), input_range[0].float(), input_range[1].float(), self.pos_enc.gauss_B.float(), x16)
        if self.feat_proj._needs_grad():   # the Fourier features carry no gradient (computed under no_grad, :22-24)
            from .train_blocks import linear_ln_train
            return linear_ln_train(None, self.feat_proj[0], self.feat_proj[1], x16=x16).view(B, L, D)
        out = torch.empty(B * L, D, dtype=torch.float32, device=coords.device)
        return self.feat_proj.run16(x16, B * L, out).view(B, L, D)


class ObjectEncoder(nn.Module):
    """Projection branch of modules/vision/object_encoder.py:14-79: Linear(Cin->D) + LayerNorm
    (+ Dropout, identity in eval).  The PointNet++ backbone and the cls head are upstream feature
    extraction, out of scope."""

    def __init__(self, cfg=None, backbone="none", input_feat_size=768, hidden_size=768, freeze_backbone=False,
                 use_projection=False, tgt_cls_num=607, pretrained=None, dropout=0.1, use_cls_head=True):
        super().__init__()
        if backbone != "none":
            raise NotImplementedError("ObjectEncoder backbone='pointnet++' (frozen tokenizer) is out of scope; "
                                      "feed its per-object features with backbone='none'")
        if use_cls_head:
            raise NotImplementedError("ObjectEncoder.use_cls_head=True is not on the Query3DUnified path "
                                      "(all shipped configs set it to False)")
        self.use_projection = use_projection
        if use_projection:
            self.input_feat_proj = LinearLN(input_feat_size, hidden_size)
        elif input_feat_size != hidden_size:
            raise AssertionError("input_feat_size should be equal to hidden_size!")
        if dropout > 0:
            self.dropout = nn.Dropout(dropout)       # object_encoder.py:39-40,75-76
        with torch.no_grad():                                     # _init_weights_bert (modules/weights.py)
            for m in self.modules():
                if isinstance(m, nn.Linear):
                    m.weight.normal_(0.0, 0.02)
                    m.bias.zero_()

    def forward(self, obj_feats, data_dict=None, emit=None, **kwargs):
        obj_embeds = self.input_feat_proj(obj_feats, emit=emit) if self.use_projection else obj_feats
        if self.training and hasattr(self, "dropout"):
            # the only torch op on this module's path: an elementwise Bernoulli mask on the producer side of the
            # decoder (§8f-1), drawn from torch's RNG exactly like the reference's nn.Dropout (object_encoder.py:75-76)
            obj_embeds = self.dropout(obj_embeds)
        return obj_embeds


class GroundHead(nn.Module):
    """modules/heads/grounding_head.py:42-55."""

    def __init__(self, cfg=None, input_size=768, hidden_size=768, dropout=0.3):
        super().__init__()
        self.og3d_head = mlp_head_params(input_size, hidden_size, 1, dropout=dropout)
        self._run = MlpHeadRunner(self.og3d_head)

    def forward(self, obj_embeds, obj_masks=None, **kwargs):
        B, N, D = obj_embeds.shape
        if self.training or (torch.is_grad_enabled() and (obj_embeds.requires_grad
                                                          or any(p.requires_grad for p in self.parameters()))):
            from .train_blocks import mlp_head_train
            logits = mlp_head_train(self.og3d_head, obj_embeds).squeeze(2)
            if obj_masks is not None:
                logits = logits.masked_fill(obj_masks.logical_not(), -float("inf"))
            return logits
        x16 = torch.empty(B * N, D, dtype=bf16, device=obj_embeds.device)
        ops.cast_bf16(obj_embeds.reshape(B * N, D).contiguous().float(), x16)
        logits = self._run(x16, B * N).view(B, N)
        if obj_masks is not None:
            logits = logits.masked_fill_(obj_masks.logical_not(), -float("inf"))
        return logits


MODULES = {"QueryMaskEncoder": QueryMaskEncoder, "QueryEncoder": QueryEncoder, "MaskHeadSegLevel": MaskHeadSegLevel,
           "ObjectEncoder": ObjectEncoder, "GroundHead": GroundHead}


def build_module_by_name(cfg):
    """modules/build.py:24-31 against this package's own name -> class table."""
    name = cfg_get(cfg, "name")
    if name not in MODULES:
        raise NotImplementedError(f"Unknown module: {name} (pq3d_b200 provides {sorted(MODULES)})")
    args = cfg_get(cfg, "args")
    kwargs = cfg_plain(args) if args is not None else {}
    return MODULES[name](cfg, **kwargs)


class Query3DUnified(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        model = cfg_get(cfg, "model")
        self.memories = list(cfg_get(model, "memories"))
        self.heads = list(cfg_get(model, "heads"))
        self.use_offline_voxel_fts = cfg_get(model, "use_offline_voxel_fts", False)
        self.use_offline_attn_mask = cfg_get(model, "use_offline_attn_mask", False)
        self.inputs = self.memories[:]
        self.pairwise_rel_type = cfg_path(model, "obj_loc.pairwise_rel_type")
        self.spatial_dim = cfg_path(model, "obj_loc.spatial_dim")
        self.num_heads = cfg_path(model, "unified_encoder.args.num_attention_heads")
        self.skip_query_encoder_mask_pred = cfg_get(model, "skip_query_encoder_mask_pred", False)
        if self.pairwise_rel_type != "center" or self.spatial_dim != 5:
            raise NotImplementedError("only pairwise_rel_type='center', spatial_dim=5 (all shipped configs)")
        self.prompt_types = ["txt", "loc"]
        for inp in self.inputs:
            if inp == "prompt":
                # txt_encoder (frozen CLIP tower + projection) is out of scope: prompt features enter
                # as data_dict['prompt_feat'] (B, T, hidden); see prompt_encoder()
                continue
            if inp == "voxel" and not self.use_offline_voxel_fts:
                raise NotImplementedError("online voxel features need the MinkowskiEngine sparse-conv backbone "
                                          "(out of scope); set model.use_offline_voxel_fts=True")
            setattr(self, inp + "_encoder", build_module_by_name(cfg_get(model, inp + "_encoder")))
        dim_loc = cfg_path(model, "obj_loc.dim_loc")
        hidden_size = cfg_get(model, "hidden_size")
        self.dim_loc, self.hidden_size = dim_loc, hidden_size
        if dim_loc > 3:
            self.coord_encoder = LinearLN(3, hidden_size)
            self.box_encoder = LinearLN(3, hidden_size)
        else:
            self.coord_encoder = CoordinateEncoder(hidden_size)
        self.unified_encoder = build_module_by_name(cfg_get(model, "unified_encoder"))
        for head in self.heads:
            if head == "generation":
                raise NotImplementedError("the T5 generation head is out of scope (needs pretrained weights)")
            setattr(self, head + "_head", build_module_by_name(cfg_get(model, head + "_head")))

    def prompt_encoder(self, data_dict):
        """model/query3d_unified.py:80-108 with the text tower factored out: `prompt_feat` (B,T,D) is
        whatever txt_encoder would have produced; returns (feat, mask True=ignore)."""
        if "prompt_type" in data_dict and "prompt" in data_dict:
            # mixed batches (:80-108): location prompts go through coord_encoder (+ box_encoder), are broadcast over
            # the T prompt slots and keep only slot 0 valid; text prompts take their precomputed features
            prompt, ptype, pad = data_dict["prompt"], data_dict["prompt_type"], data_dict["prompt_pad_masks"]
            feat = torch.zeros(tuple(prompt.shape) + (self.hidden_size,), device=prompt.device)
            txt, loc = ptype == PROMPT_TYPE_TXT, ptype == PROMPT_TYPE_LOC
            if bool(((~txt) & (~loc)).any()):
                raise NotImplementedError("prompt types other than TXT (1) and LOC (3) are not implemented")
            if bool(txt.any()):
                if "prompt_feat" not in data_dict:
                    raise NotImplementedError("pq3d_b200.Query3DUnified takes precomputed text-prompt features in "
                                              "data_dict['prompt_feat'] (the CLIP text tower is out of scope)")
                feat[txt] = data_dict["prompt_feat"][txt].to(feat.dtype)
            if bool(loc.any()):
                lp = prompt[loc][:, :self.dim_loc].float()
                if self.dim_loc > 3:
                    f = self.coord_encoder(lp[:, :3]).unsqueeze(1) + self.box_encoder(lp[:, 3:6]).unsqueeze(1)
                else:
                    f = self.coord_encoder(lp[:, :3].unsqueeze(1),
                                           input_range=[data_dict["coord_min"][loc], data_dict["coord_max"][loc]])
                feat[loc] = f.to(feat.dtype)                       # (n, 1, D) broadcast over the T slots, as the reference
                m = pad[loc]
                m[:, 1:] = False
                pad[loc] = m
            return feat, pad.logical_not()
        if "prompt_feat" not in data_dict:
            raise NotImplementedError("pq3d_b200.Query3DUnified takes precomputed prompt features in "
                                      "data_dict['prompt_feat'] (the CLIP text tower is out of scope)")
        return data_dict["prompt_feat"], data_dict["prompt_pad_masks"].logical_not()

    # ---- whole-model CUDA graph (inference serving loops) ---------------------------------------------------------
    def _graph_signature(self, data_dict):
        """(hashable signature, tensors): every CUDA tensor of the batch by address / shape / strides / dtype, the CUDA
        stream, and the switches the captured launch sequence depends on."""
        # host tensors (labels, ids) are not part of the signature: nothing on this path can read them
        items = [(k, v) for k, v in sorted(data_dict.items(), key=lambda kv: str(kv[0]))
                 if isinstance(v, torch.Tensor) and v.is_cuda]
        if not items:
            return None, None
        dev = items[0][1].device
        sig = (torch.cuda.current_stream(dev).cuda_stream, os.environ.get("PQ3D_PREINGEST", "1"),
               tuple((k, v.data_ptr(), tuple(v.shape), v.stride(), v.dtype) for k, v in items))
        return sig, [v for _, v in items]

    def _forward_graphed(self, data_dict):
        """A serving loop that hands in the SAME device tensors again (staging buffers refilled in place) gets the whole
        forward — coordinate encoder, ObjectEncoder projections, pairwise geometry, decoder, grounding head: ~100 kernel
        launches and as many torch ops, 1.3 ms of host time against 0.9 ms on the device — as ONE graph launch.  A batch
        signature is captured the second time it is seen with the very same tensor objects (addresses alone could be
        allocator re-use); the replay reads the buffers' current contents and the results are returned as fresh tensors.
        Returns None when the call is not eligible (training, autograd, mixed prompt batches whose control flow depends
        on device values, a mask head, layer taps) and the eager path runs."""
        enc = self.unified_encoder
        if (not getattr(self, "use_cuda_graph", True) or not getattr(enc, "use_cuda_graph", False) or self.training
                or torch.is_grad_enabled() or "prompt_type" in data_dict or hasattr(self, "mask_head")
                or any(h != "ground" for h in self.heads) or getattr(enc, "layer_taps", None) is not None
                or type(enc) is not QueryMaskEncoder or not torch.cuda.is_available()
                or torch.cuda.is_current_stream_capturing()):
            return None
        sig, tensors = self._graph_signature(data_dict)
        if sig is None:
            return None
        graphs = self.__dict__.setdefault("_graphs", {})
        # the captured launches read the kernels' operand copies of the parameters (bf16 weights, packed decoder tables):
        # any in-place update (optimizer step, load_state_dict -> version counters) or move (.to(): addresses) since the
        # capture re-creates those copies, so every graph is dropped and captured afresh
        plist = self.__dict__.get("_graph_params")
        if plist is None:
            plist = self.__dict__["_graph_params"] = list(self.parameters()) + list(self.buffers())
        wkey = (tuple(p._version for p in plist), tuple(p.data_ptr() for p in plist))
        if self.__dict__.get("_graph_wkey") != wkey:
            graphs.clear()
            self.__dict__["_graph_wkey"] = wkey
        ent = graphs.get(sig)
        if ent is not None and ent.get("graph") is None and not all(r() is t for r, t in zip(ent["refs"], tensors)):
            ent = None
        if ent is None:
            if len(graphs) >= 8:
                graphs.pop(next(iter(graphs)))
            graphs[sig] = {"refs": [weakref.ref(t) for t in tensors]}
            return None                                          # first sighting: eager (allocates every workspace)
        if ent.get("failed"):
            return None
        if ent.get("graph") is None:
            torch.cuda.synchronize()
            before = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            dd = dict(data_dict)
            enc._stream_key_override = sig[0]                    # capture runs on torch's side stream: keep the caller's workspace
            try:
                with torch.cuda.graph(g):
                    self._forward_eager(dd)
            except RuntimeError as e:
                # a configuration whose forward cannot be recorded (an op that synchronises): stay on the eager path for
                # this signature and say so once — never a silent change of results, the eager forward is the reference
                import warnings
                warnings.warn(f"pq3d_b200.Query3DUnified: whole-model CUDA graph capture failed ({e}); this batch "
                              "signature keeps the eager forward")
                ent["failed"] = True
                ops.LAUNCHES = before
                torch.cuda.synchronize()
                return None
            finally:
                enc._stream_key_override = None
            outs = {k: v for k, v in dd.items()
                    if isinstance(v, torch.Tensor) and data_dict.get(k) is not v and k != "ground_label"}
            ent.update(graph=g, outs=outs, launches=ops.LAUNCHES - before, keep=tensors)
            ops.LAUNCHES = before
        ent["graph"].replay()
        ops._count(ent["launches"])
        fresh = {}
        for k, v in ent["outs"].items():
            if id(v) not in fresh:
                fresh[id(v)] = v.clone()
            data_dict[k] = fresh[id(v)]
        data_dict["ground_label"] = data_dict.get("tgt_object_id")
        return data_dict

    def forward(self, data_dict):
        if self.training or torch.is_grad_enabled():
            # fused optimizers update parameters without bumping version counters: a training-mode call invalidates
            # the inference graphs outright
            self.__dict__.get("_graphs", {}).clear()
            return self._forward_eager(data_dict)
        out = self._forward_graphed(data_dict)
        return out if out is not None else self._forward_eager(data_dict)

    def _forward_eager(self, data_dict):
        input_dict = {}
        mask = data_dict["query_pad_masks"].logical_not()
        query_locs = data_dict["query_locs"][:, :, :self.dim_loc]
        coord_min, coord_max = data_dict["coord_min"], data_dict["coord_max"]
        fts_locs = data_dict["seg_center"]
        if self.dim_loc > 3:
            query_pos = self.coord_encoder(query_locs[:, :, :3]) + self.box_encoder(query_locs[:, :, 3:6])
            box_pos = self.box_encoder(fts_locs[:, :, 3:6])
            # the reference adds box_encoder(fts_locs[..., 3:6]) twice (query3d_unified.py:128,131-132)
            fts_pos = self.coord_encoder(fts_locs[:, :, :3]) + box_pos
            fts_pos += box_pos
        else:
            query_pos = self.coord_encoder(query_locs[:, :, :3], input_range=[coord_min, coord_max])
            fts_pos = self.coord_encoder(fts_locs[:, :, :3], input_range=[coord_min, coord_max])
        input_dict["query"] = (torch.zeros_like(query_pos), mask, query_pos)
        # §8f-1: in inference the producers' LayerNorm epilogue writes the decoder's bf16 K / V operands of the scene
        # memories directly (xv = bf16(feat), xk = bf16(feat + fts_pos)), stacked the way the grouped K / V^T projection
        # reads them — the decoder then skips its ingest pass over the fp32 tables
        emit_of = self._plan_preingest(fts_pos)
        for inp in self.inputs:
            feat, mask, pos = None, None, None
            if inp == "prompt":
                feat, mask = self.prompt_encoder(data_dict)
            elif inp == "mv":
                feat = self.mv_encoder(obj_feats=data_dict["mv_seg_fts"], emit=emit_of.get("mv"))
                mask = data_dict["mv_seg_pad_masks"].logical_not()
                pos = fts_pos
            elif inp == "pc":
                feat = self.pc_encoder(obj_feats=data_dict["pc_seg_fts"], emit=emit_of.get("pc"))
                mask = data_dict["pc_seg_pad_masks"].logical_not()
                pos = fts_pos
            elif inp == "voxel":
                feat = self.voxel_encoder(data_dict["voxel_seg_fts"], emit=emit_of.get("voxel"))
                mask = data_dict["voxel_seg_pad_masks"].logical_not()
                pos = fts_pos
            else:
                raise NotImplementedError(f"Unknow input type: {inp}")
            input_dict[inp] = [feat, mask, pos]
        offline_attn_masks = data_dict["offline_attn_mask"] if self.use_offline_attn_mask else None
        seg_fts_for_match = []
        for inp in self.inputs:
            if inp in ("voxel", "mv", "pc"):
                feats = copy(input_dict[inp][:])
                if isinstance(feats[0], list):
                    assert inp == "voxel"
                    feats[0] = feats[0][-1]
                seg_fts_for_match.append(feats)
        if hasattr(self, "mask_head"):
            mask_head_partial = partial(self.mask_head, seg_fts_for_match=seg_fts_for_match,
                                        seg_masks=data_dict["seg_pad_masks"].logical_not(),
                                        offline_attn_masks=offline_attn_masks,
                                        skip_prediction=self.skip_query_encoder_mask_pred)
        else:
            mask_head_partial = None
        pairwise_locs = ops.pairwise_locs(query_locs[:, :, :3].float()) if self.unified_encoder.spatial_selfattn else None

        query, predictions_class, predictions_mask = self.unified_encoder(input_dict, pairwise_locs, mask_head_partial)

        for head in self.heads:
            if head == "ground":
                logits = self.ground_head(query, data_dict["query_pad_masks"])
                data_dict["ground_logits"] = logits
                data_dict["og3d_logits"] = logits
                data_dict["ground_label"] = data_dict.get("tgt_object_id")
            elif head == "mask":
                if self.skip_query_encoder_mask_pred:
                    mask_head_partial = partial(self.mask_head, seg_fts_for_match=seg_fts_for_match,
                                                seg_masks=data_dict["seg_pad_masks"].logical_not(),
                                                offline_attn_masks=offline_attn_masks, skip_prediction=False)
                    predictions_class, predictions_mask = [], []
                pred_logits, pred_masks, _ = mask_head_partial(query=query)
                predictions_class.append(pred_logits)
                predictions_mask.append(pred_masks)
                data_dict["predictions_class"] = predictions_class
                data_dict["predictions_mask"] = predictions_mask
            else:
                raise NotImplementedError(f"Unknow head type: {head}")
        return data_dict

    def _plan_preingest(self, fts_pos: torch.Tensor) -> dict:
        """{memory: (xv16, xk16, pos32)} for the scene memories whose encoder can emit the decoder's operands, or {}."""
        enc = self.unified_encoder
        if hasattr(enc, "preingested"):
            enc.preingested = None
        if (self.training or torch.is_grad_enabled() or not isinstance(enc, QueryMaskEncoder)
                or os.environ.get("PQ3D_PREINGEST", "1") == "0"):
            return {}
        mems = [m for m in enc._active() if m != "prompt"]
        B, S, D = fts_pos.shape
        ok = (len(mems) >= 2 and S % 8 == 0 and fts_pos.dtype == torch.float32 and fts_pos.is_contiguous()
              and all(m in self.inputs and getattr(getattr(self, m + "_encoder"), "use_projection", False) for m in mems)
              and all(getattr(self, m + "_encoder").input_feat_proj[0].out_features == D for m in mems))
        if not ok:
            return {}
        key = (B, S, D, len(mems), str(fts_pos.device))
        if getattr(self, "_pre_key", None) != key:
            self._pre_key = key
            self._pre_buf = (torch.empty(len(mems) * B * S, D, dtype=bf16, device=fts_pos.device),
                             torch.empty(len(mems) * B * S, D, dtype=bf16, device=fts_pos.device))
        xk_all, xv_all = self._pre_buf
        pos2 = fts_pos.view(B * S, D)
        plan = {m: (xv_all[j * B * S:(j + 1) * B * S], xk_all[j * B * S:(j + 1) * B * S], pos2) for j, m in enumerate(mems)}
        enc.preingested = dict(mems=tuple(mems), xk=xk_all, xv=xv_all, pos=fts_pos, S=S, B=B)
        return plan

    def get_opt_params(self):
        """model/query3d_unified.py:224-238 with optim/utils.py:1-18 inlined: per-submodule groups,
        weight decay 0.01 except names containing 'bias' / 'LayerNorm.bias' / 'LayerNorm.weight'."""
        default_lr = cfg_path(self.cfg, "solver.lr")
        groups = []
        no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
        for name, module in self._modules.items():
            lr = cfg_get(cfg_path(self.cfg, "model." + name), "lr", None) or default_lr
            if lr != default_lr:
                print(f"Change lr from default {default_lr} to {lr} for {name} module.")
            decay, nodecay = [], []
            for n, p in module.named_parameters():
                if not p.requires_grad:
                    continue
                (nodecay if any(nd in n for nd in no_decay) else decay).append(p)
            groups += [{"params": decay, "name": name, "weight_decay": 0.01, "lr": lr},
                       {"params": nodecay, "name": name, "weight_decay": 0.0, "lr": lr}]
        optimized = [p for g in groups for p in g["params"]]
        assert len(optimized) == len(list(self.parameters())), "Some parameters are not optimized!"
        return groups
