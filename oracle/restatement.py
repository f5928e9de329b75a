"""TEST INFRASTRUCTURE — the CPU/torch oracle for the PQ3D promptable-query-decoder hot path.

A from-first-principles, *functional* restatement (plain torch ops over a flat state_dict) of the
reference algorithm.  It is the checker for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it,
never the product package `pq3d_b200`.

Parity pin: the reference holds no golden vectors / known-answer tests for this path (SURVEY.md
§4, §8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in this container
by `oracle/ref_loader.py` (`tests/test_oracle_vs_reference.py`, fp32, <=1e-5) and against the
committed fixtures under `tests/golden/` which `oracle/make_golden.py` generated from the real
reference modules.

The op sequence inside `mha()` deliberately mirrors torch's
`F.multi_head_attention_forward(need_weights=True)` branch (linear -> baddbmm/bmm -> softmax ->
bmm -> linear) so that under `torch.autocast('cuda', torch.bfloat16)` it rounds at exactly the
points the reference does — that is the "reference GPU path" timed beside the kernels.

Every function cites the reference file:line it follows (paths relative to /root/reference, or
`torch/` for the installed PyTorch).  Masks: True = ignore everywhere in here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass
class DecoderCfg:
    """kwargs of QueryMaskEncoder.__init__ (modules/grounding/query_encoder.py:53-54)."""
    memories: Sequence[str] = ()
    memory_dropout: float = 0.0
    hidden_size: int = 768
    num_attention_heads: int = 12
    num_layers: int = 4
    spatial_selfattn: bool = False
    structure: str = "sequential"
    drop_memories_test: Sequence[str] = field(default_factory=list)
    use_self_mask: bool = False
    num_blocks: int = 1
    # Training mode (module.train()): an object supplying the random decisions the reference draws from torch's RNG, so
    # a checker can replay the exact masks another implementation used.  None = eval mode.  Methods:
    #   sublayer(layer, kind, x)  -> dropout(x): kind = ('ca', memory name) | 'sa' | 'ffn'   (nn.Dropout on tgt2)
    #   probs(layer, kind, P)     -> dropout(P) on attention probabilities (B*H, L, S[+1]); kind = ('ca', name) | 'sa'
    #   hidden(layer, h)          -> dropout(h) on the FFN hidden activations
    #   memory_keep(layer, memories, B) -> bool (B, len(memories)) keep mask BEFORE the keep-all fix-up, or None
    train: Optional[object] = None


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def layer_norm(x: Tensor, sd: SD, prefix: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def mha(q_in: Tensor, k_in: Tensor, v_in: Tensor, sd: SD, prefix: str, nhead: int,
        key_padding_mask: Optional[Tensor], attn_mask: Optional[Tensor], add_zero_attn: bool,
        prob_dropout: Optional[Callable] = None) -> Tensor:
    """nn.MultiheadAttention(batch_first=True) forward, eval mode, need_weights=True branch.

    torch/nn/functional.py:5867-5873 (separate q/k/v in-projections from the packed weight),
    :6585-6602 (add_zero_attn: one zero key/value appended AFTER projection, masks padded with
    False), :6608-6620 (key padding merged into the float mask), :6630-6654 (scaled q, baddbmm,
    softmax, bmm, out_proj).  Inputs are (B, L, E) / (B, S, E); bool masks, True = ignore.
    """
    B, L, E = q_in.shape
    S = k_in.shape[1]
    dh = E // nhead
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = F.linear(q_in, w[:E], b[:E])
    k = F.linear(k_in, w[E:2 * E], b[E:2 * E])
    v = F.linear(v_in, w[2 * E:], b[2 * E:])
    # (B, L, H, dh) -> (B*H, L, dh); row b*H + h, as the reference's view/transpose produces
    q = q.view(B, L, nhead, dh).transpose(1, 2).reshape(B * nhead, L, dh)
    k = k.view(B, S, nhead, dh).transpose(1, 2).reshape(B * nhead, S, dh)
    v = v.view(B, S, nhead, dh).transpose(1, 2).reshape(B * nhead, S, dh)

    fmask = None
    if attn_mask is not None:
        assert attn_mask.dtype == torch.bool and attn_mask.shape == (B * nhead, L, S)
        fmask = torch.zeros(attn_mask.shape, dtype=q_in.dtype, device=q_in.device).masked_fill_(attn_mask, float("-inf"))
    kpm = None
    if key_padding_mask is not None:
        assert key_padding_mask.dtype == torch.bool and key_padding_mask.shape == (B, S)
        kpm = torch.zeros(key_padding_mask.shape, dtype=q_in.dtype, device=q_in.device).masked_fill_(key_padding_mask, float("-inf"))
    if add_zero_attn:
        k = torch.cat([k, k.new_zeros(B * nhead, 1, dh)], dim=1)
        v = torch.cat([v, v.new_zeros(B * nhead, 1, dh)], dim=1)
        if fmask is not None:
            fmask = F.pad(fmask, (0, 1))
        if kpm is not None:
            kpm = F.pad(kpm, (0, 1))
    S2 = k.shape[1]
    if kpm is not None:
        kpm = kpm.view(B, 1, 1, S2).expand(-1, nhead, -1, -1).reshape(B * nhead, 1, S2)
        fmask = kpm if fmask is None else fmask + kpm

    q_scaled = q * math.sqrt(1.0 / float(dh))
    if fmask is not None:
        scores = torch.baddbmm(fmask, q_scaled, k.transpose(-2, -1))
    else:
        scores = torch.bmm(q_scaled, k.transpose(-2, -1))
    probs = F.softmax(scores, dim=-1)
    if prob_dropout is not None:                                # training: dropout(attn_output_weights), :6645-6646
        probs = prob_dropout(probs)
    out = torch.bmm(probs, v)                                   # (B*H, L, dh)
    out = out.view(B, nhead, L, dh).transpose(1, 2).reshape(B, L, E)
    return F.linear(out, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def cross_attention_layer(tgt: Tensor, memory: Tensor, sd: SD, prefix: str, nhead: int,
                          attn_mask: Optional[Tensor], key_padding_mask: Optional[Tensor],
                          pos: Optional[Tensor], query_pos: Optional[Tensor], prob_dropout: Optional[Callable] = None,
                          dropout: Optional[Callable] = None) -> Tensor:
    """CrossAttentionLayer.forward_post (modules/grounding/query_encoder.py:288-307);
    MHA built with add_zero_attn=True and dropout=p (:268-270); `tgt + dropout(tgt2)` (:304)."""
    q_in = tgt if query_pos is None else tgt + query_pos
    k_in = memory if pos is None else memory + pos
    upd = mha(q_in, k_in, memory, sd, prefix + "multihead_attn.", nhead, key_padding_mask, attn_mask, True, prob_dropout)
    if dropout is not None:
        upd = dropout(upd)
    return layer_norm(tgt + upd, sd, prefix + "norm.")


def self_attention_layer(tgt: Tensor, sd: SD, prefix: str, nhead: int,
                         key_padding_mask: Optional[Tensor], query_pos: Optional[Tensor],
                         prob_dropout: Optional[Callable] = None, dropout: Optional[Callable] = None) -> Tensor:
    """SelfAttentionLayer.forward_post (query_encoder.py:213-227): q = k = tgt+pos, v = tgt."""
    x = tgt if query_pos is None else tgt + query_pos
    upd = mha(x, x, tgt, sd, prefix + "self_attn.", nhead, key_padding_mask, None, False, prob_dropout)
    if dropout is not None:
        upd = dropout(upd)
    return layer_norm(tgt + upd, sd, prefix + "norm.")


def spatial_mha(x: Tensor, v_in: Tensor, pairwise_locs: Tensor, sd: SD, prefix: str, nhead: int,
                key_padding_mask: Optional[Tensor]) -> Tensor:
    """MultiHeadAttentionSpatial.forward, fusion 'mul', spatial_multihead=True
    (modules/layers/transformers.py:189-240)."""
    B, L, E = x.shape
    dh = E // nhead

    def heads(t):  # 'b l (head k) -> head b l k'
        return t.view(B, L, nhead, dh).permute(2, 0, 1, 3)

    q = heads(F.linear(x, sd[prefix + "w_qs.weight"], sd[prefix + "w_qs.bias"]))
    k = heads(F.linear(x, sd[prefix + "w_ks.weight"], sd[prefix + "w_ks.bias"]))
    v = heads(F.linear(v_in, sd[prefix + "w_vs.weight"], sd[prefix + "w_vs.bias"]))
    attn = torch.einsum("hblk,hbtk->hblt", q, k) / math.sqrt(dh)                      # :193
    loc = F.linear(pairwise_locs, sd[prefix + "pairwise_loc_fc.weight"], sd[prefix + "pairwise_loc_fc.bias"])
    loc = F.relu(loc.permute(3, 0, 1, 2))                                            # (H,B,L,T) :196-199
    if key_padding_mask is not None:                                                 # :220-226
        m = key_padding_mask.view(1, B, 1, L).expand(nhead, B, L, L)
        attn = attn.masked_fill(m, float("-inf"))
        loc = loc.masked_fill(m, 0)
    fused = torch.log(torch.clamp(loc, min=1e-6)) + attn                             # :231-232
    fused = torch.softmax(fused, 3)
    assert not torch.isnan(fused).any()                                              # :235
    out = torch.einsum("hblt,hbtv->hblv", fused, v)
    out = out.permute(1, 2, 0, 3).reshape(B, L, E)
    return F.linear(out, sd[prefix + "fc.weight"], sd[prefix + "fc.bias"])


def spatial_self_attention_layer(tgt: Tensor, sd: SD, prefix: str, nhead: int,
                                 key_padding_mask: Optional[Tensor], query_pos: Optional[Tensor],
                                 pairwise_locs: Tensor, dropout: Optional[Callable] = None) -> Tensor:
    """SpatialSelfAttentionLayer.forward_post (query_encoder.py:438-452).  MultiHeadAttentionSpatial never applies its
    `dropout` argument (transformers.py:158-240), so only the sublayer dropout (:449) exists in training."""
    x = tgt if query_pos is None else tgt + query_pos
    upd = spatial_mha(x, tgt, pairwise_locs, sd, prefix + "self_attn.", nhead, key_padding_mask)
    if dropout is not None:
        upd = dropout(upd)
    return layer_norm(tgt + upd, sd, prefix + "norm.")


def ffn_layer(tgt: Tensor, sd: SD, prefix: str, hidden_dropout: Optional[Callable] = None,
              dropout: Optional[Callable] = None) -> Tensor:
    """FFNLayer.forward_post, relu (query_encoder.py:384-388): linear2(dropout(relu(linear1 x))), tgt + dropout(tgt2)."""
    h = F.relu(F.linear(tgt, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"]))
    if hidden_dropout is not None:
        h = hidden_dropout(h)
    upd = F.linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    if dropout is not None:
        upd = dropout(upd)
    return layer_norm(tgt + upd, sd, prefix + "norm.")


# --------------------------------------------------------------------------------------------
# one decoder layer / the stacked decoder
# --------------------------------------------------------------------------------------------
def query_encoder_layer(query: Tensor, input_dict: dict, pairwise_locs: Optional[Tensor], sd: SD,
                        prefix: str, cfg: DecoderCfg, layer: int = 0) -> Tensor:
    """QueryEncoderLayer.forward (query_encoder.py:114-181); eval mode unless cfg.train supplies the random draws."""
    H = cfg.num_attention_heads
    _, query_masks, query_pos = input_dict["query"]
    mem_index = {m: j for j, m in enumerate(cfg.memories)}            # memory2ca, :105
    tr = cfg.train

    def hook(name, *key):
        return None if tr is None else (lambda x: getattr(tr, name)(layer, *key, x))

    def one_ca(q, memory):
        feat, mask, pos = input_dict[memory]
        kpm, am = (mask, None) if mask.ndim == 2 else (None, mask)     # :121-126
        return cross_attention_layer(q, feat, sd, f"{prefix}cross_attn_list.{mem_index[memory]}.", H,
                                     am, kpm, pos, query_pos, hook("probs", ("ca", memory)),
                                     hook("sublayer", ("ca", memory)))

    def sequential_ca(q, memories):                                    # :117-129
        for m in memories:
            q = one_ca(q, m)
        return q

    def parallel_ca(q, memories):                                      # :131-154
        assert "prompt" not in memories
        stack = torch.stack([one_ca(q, m) for m in memories], dim=1)
        keep = None
        if tr is not None and cfg.memory_dropout > 0.0:               # :145-152
            keep = tr.memory_keep(layer, list(memories), q.shape[0])
        if keep is None:
            return stack.mean(dim=1)
        n_left = keep.sum(dim=1)
        keep = torch.logical_or(keep, n_left.unsqueeze(-1) == 0)
        n_left = keep.sum(dim=1)
        return (stack * keep.unsqueeze(-1).unsqueeze(-1)).sum(dim=1) / n_left.unsqueeze(-1).unsqueeze(-1).float()

    if tr is not None:
        memories = list(cfg.memories)                                  # training: every memory, :156
    else:
        memories = [m for m in cfg.memories if m not in cfg.drop_memories_test]   # :156
    if cfg.structure == "sequential":
        query = sequential_ca(query, memories)
    elif cfg.structure == "parallel":
        query = parallel_ca(query, memories)
    elif cfg.structure == "mixed":                                     # :162-165
        query = parallel_ca(query, [m for m in memories if m != "prompt"])
        query = sequential_ca(query, ["prompt"])
    elif cfg.structure == "gate":                                      # :166-170
        prompt = sequential_ca(query, ["prompt"])
        gate = torch.sigmoid(F.linear(prompt, sd[prefix + "gate_proj.weight"], sd[prefix + "gate_proj.bias"]))
        update = parallel_ca(query, [m for m in cfg.memories if m != "prompt"])
        query = (1.0 - gate) * query + gate * update
    else:
        raise NotImplementedError(f"Unknow structure type: {cfg.structure}")

    if cfg.spatial_selfattn:                                           # :174-178
        query = spatial_self_attention_layer(query, sd, prefix + "self_attn.", H, query_masks, query_pos, pairwise_locs,
                                             hook("sublayer", "sa"))
    else:
        query = self_attention_layer(query, sd, prefix + "self_attn.", H, query_masks, query_pos, hook("probs", "sa"),
                                     hook("sublayer", "sa"))
    return ffn_layer(query, sd, prefix + "ffn.", hook("hidden"), hook("sublayer", "ffn"))        # :179


def query_mask_encoder(sd: SD, cfg: DecoderCfg, input_dict: dict, pairwise_locs: Optional[Tensor],
                       mask_head: Optional[Callable] = None, prefix: str = ""):
    """QueryMaskEncoder.forward (query_encoder.py:69-94).  Mutates input_dict like the reference."""
    predictions_class, predictions_mask = [], []
    query = input_dict["query"][0]
    voxel_feat = input_dict["voxel"][0] if "voxel" in input_dict else None
    attn_mask = None
    for _block in range(cfg.num_blocks):
        for i in range(cfg.num_layers):
            if mask_head is not None:
                output_class, outputs_mask, attn_mask = mask_head(query)
                predictions_class.append(output_class)
                predictions_mask.append(outputs_mask)
            if cfg.use_self_mask:
                attn_mask[attn_mask.all(-1)] = False                      # :83
                attn_mask = attn_mask.repeat_interleave(cfg.num_attention_heads, 0)
                for memory in input_dict.keys():
                    if memory in ("query", "prompt"):
                        continue
                    input_dict[memory][1] = attn_mask                      # :85-88 (overwrite)
            if isinstance(voxel_feat, list):
                input_dict["voxel"][0] = voxel_feat[i]                     # :90-91
            query = query_encoder_layer(query, input_dict, pairwise_locs, sd,
                                        f"{prefix}unified_encoder.{i}.", cfg, layer=i)
    return query, predictions_class, predictions_mask


# --------------------------------------------------------------------------------------------
# mask head (in-loop consumer/producer of the attention mask)
# --------------------------------------------------------------------------------------------
def mlp_head(x: Tensor, sd: SD, prefix: str, dropout: Optional[Callable] = None) -> Tensor:
    """get_mlp_head: Linear-ReLU-LayerNorm(eps 1e-12)-Dropout-Linear (modules/utils.py:18-25); `dropout` (training
    only) is applied where the nn.Dropout sits."""
    h = F.relu(F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    h = F.layer_norm(h, (h.shape[-1],), sd[prefix + "2.weight"], sd[prefix + "2.bias"], 1e-12)
    if dropout is not None:
        h = dropout(h)
    return F.linear(h, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def mask_head_seg_level(query: Tensor, sd: SD, prefix: str, seg_fts_for_match: list, seg_masks: Tensor,
                        filter_out_classes=None, offline_attn_masks: Optional[Tensor] = None,
                        skip_prediction: bool = False, cls_dropout: Optional[Callable] = None):
    """MaskHeadSegLevel.forward + MaskPredictionLayer (modules/heads/mask_head.py:24-57)."""
    if skip_prediction:
        return None, None, offline_attn_masks
    cls_logits = mlp_head(query, sd, prefix + "cls_head.", cls_dropout)
    # NB the reference indexes unconditionally (mask_head.py:28); with filter_out_classes=None the
    # index `[..., None]` addresses every class, so all logits become -inf.  Reproduced as is.
    cls_logits[..., filter_out_classes] = float("-inf")
    logits_sum, valid_sum = 0, 0
    for j, (feat, mask, _pos) in enumerate(seg_fts_for_match):
        qp = F.linear(query, sd[f"{prefix}mask_pred_list.{j}.q_proj.weight"], sd[f"{prefix}mask_pred_list.{j}.q_proj.bias"])
        kp = F.linear(feat, sd[f"{prefix}mask_pred_list.{j}.k_proj.weight"])
        logits = torch.einsum("bld,bmd->blm", kp, qp)                     # (B, S, N)
        valid = mask[..., None].logical_not()
        logits_sum = logits_sum + logits * valid
        valid_sum = valid_sum + valid
    mask_logits = logits_sum / (valid_sum + 1e-8)
    mask_logits[seg_masks] = -1e6
    if offline_attn_masks is not None:
        attn_mask = offline_attn_masks
    else:
        attn_mask = mask_logits.sigmoid().permute(0, 2, 1).detach() < 0.5
    return cls_logits, mask_logits, attn_mask


# --------------------------------------------------------------------------------------------
# geometry / positional producers (cheap, fp32)
# --------------------------------------------------------------------------------------------
def calc_pairwise_locs(centers: Tensor, eps: float = 1e-10) -> Tensor:
    """calc_pairwise_locs(pairwise_rel_type='center', spatial_dist_norm=True, spatial_dim=5)
    (modules/utils.py:38-68): (B, N, 3) -> (B, N, N, 5)."""
    d = centers[:, :, None, :] - centers[:, None, :, :]
    dist = torch.sqrt((d ** 2).sum(3) + eps)
    mx = dist.flatten(1).max(dim=1)[0]
    norm = dist / mx[:, None, None]
    dist2d = torch.sqrt((d[..., :2] ** 2).sum(3) + eps)
    return torch.stack([norm, d[..., 2] / dist, dist2d / dist, d[..., 1] / dist2d, d[..., 0] / dist2d], dim=3)


def fourier_pos(xyz: Tensor, gauss_B: Tensor, coord_min: Tensor, coord_max: Tensor) -> Tensor:
    """PositionEmbeddingCoordsSine(pos_type='fourier', normalize=True).forward, returned already
    permuted to (B, L, d_pos) (modules/third_party/mask3d/position_embedding.py:13-43,127-156;
    model/query3d_unified.py:22-24, fp32 / no autocast)."""
    with torch.autocast(device_type=xyz.device.type, enabled=False):
        x = xyz.float()
        src_diff = coord_max[:, None, :] - coord_min[:, None, :]
        x = ((x - coord_min[:, None, :]) * 1.0) / src_diff + 0.0
        x = x * (2 * math.pi)
        proj = (x.reshape(-1, 3) @ gauss_B).view(x.shape[0], x.shape[1], -1)
        return torch.cat([proj.sin(), proj.cos()], dim=2)


def coordinate_encoder(xyz: Tensor, sd: SD, prefix: str, coord_min: Tensor, coord_max: Tensor) -> Tensor:
    """CoordinateEncoder.forward (model/query3d_unified.py:15-27): fourier -> Linear -> LN."""
    pos = fourier_pos(xyz, sd[prefix + "pos_enc.gauss_B"], coord_min, coord_max)
    pos = F.linear(pos, sd[prefix + "feat_proj.0.weight"], sd[prefix + "feat_proj.0.bias"])
    return F.layer_norm(pos, (pos.shape[-1],), sd[prefix + "feat_proj.1.weight"], sd[prefix + "feat_proj.1.bias"])


def linear_ln(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """nn.Sequential(Linear, LayerNorm) — ObjectEncoder.input_feat_proj
    (modules/vision/object_encoder.py:33-38,73) and the dim_loc>3 coord/box encoders
    (model/query3d_unified.py:62-69)."""
    h = F.linear(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"])
    return F.layer_norm(h, (h.shape[-1],), sd[prefix + "1.weight"], sd[prefix + "1.bias"])


# --------------------------------------------------------------------------------------------
# the model boundary
# --------------------------------------------------------------------------------------------
@dataclass
class ModelCfg:
    """The cfg.model.* keys Query3DUnified reads (model/query3d_unified.py:31-78)."""
    memories: Sequence[str]
    decoder: DecoderCfg
    dim_loc: int = 3
    heads: Sequence[str] = ("mask",)
    use_offline_voxel_fts: bool = True
    skip_query_encoder_mask_pred: bool = False
    filter_out_classes: Optional[List[int]] = None
    memories_for_match: Sequence[str] = ()
    projected_memories: Sequence[str] = ("mv", "pc", "voxel")   # ObjectEncoder use_projection=True
    # Training mode of the modules around the decoder (the decoder's own draws come from decoder.train): an object with
    #   obj_dropout(memory name, x)  -> dropout(x) after ObjectEncoder's projection (modules/vision/object_encoder.py:75-76)
    #   head_dropout(head name, h)   -> the Dropout inside get_mlp_head ('mask' cls head / 'ground' head)
    # None = eval mode.
    train: Optional[object] = None


def prompt_encoder(sd: SD, cfg: "ModelCfg", data_dict: dict):
    """Query3DUnified.prompt_encoder (model/query3d_unified.py:80-108) with the text tower factored out: rows of type
    TXT (PromptType.TXT = 1, data/datasets/constant.py:628-631) take `data_dict['prompt_feat']`, rows of type LOC (= 3)
    are encoded from their location by the coordinate (+ box) encoder, broadcast over the T slots, and keep only slot 0
    valid (`mask[:, 1:] = False`, written back into data_dict['prompt_pad_masks'] like the reference does).
    Returns (feat (B, T, D), mask with True = ignore)."""
    if "prompt_type" not in data_dict or "prompt" not in data_dict:
        return data_dict["prompt_feat"], data_dict["prompt_pad_masks"].logical_not()
    prompt, ptype, pad = data_dict["prompt"], data_dict["prompt_type"], data_dict["prompt_pad_masks"]
    D = sd["coord_encoder.feat_proj.0.weight" if cfg.dim_loc <= 3 else "coord_encoder.0.weight"].shape[0]
    feat = torch.zeros(tuple(prompt.shape) + (D,), device=prompt.device)
    txt, loc = ptype == 1, ptype == 3
    if bool(txt.any()):
        feat[txt] = data_dict["prompt_feat"][txt].to(feat.dtype)
    if bool(loc.any()):
        lp = prompt[loc][:, :cfg.dim_loc].float()
        if cfg.dim_loc > 3:
            f = linear_ln(lp[:, :3], sd, "coord_encoder.").unsqueeze(1) + linear_ln(lp[:, 3:6], sd, "box_encoder.").unsqueeze(1)
        else:
            f = coordinate_encoder(lp[:, :3].unsqueeze(1), sd, "coord_encoder.", data_dict["coord_min"][loc],
                                   data_dict["coord_max"][loc])
        feat[loc] = f.to(feat.dtype)          # (CPU autocast leaves LayerNorm outputs in bf16; CUDA autocast does not)
        m = pad[loc]
        m[:, 1:] = False
        pad[loc] = m
    return feat, pad.logical_not()


def query3d_unified_forward(sd: SD, cfg: ModelCfg, data_dict: dict) -> dict:
    """Query3DUnified.forward, eval mode (model/query3d_unified.py:110-222), restricted to the
    in-scope producers: offline voxel features (ObjectEncoder projection, or an already projected
    multi-scale list passed as data_dict['voxel_seg_fts_multiscale']), mv / pc ObjectEncoder
    projections, prompt features given as data_dict['prompt_feat'] (CLIP tower is out of scope),
    and the 'mask' / 'ground' heads."""
    input_dict = {}
    qmask = data_dict["query_pad_masks"].logical_not()                         # :113
    query_locs = data_dict["query_locs"][:, :, :cfg.dim_loc]
    cmin, cmax = data_dict["coord_min"], data_dict["coord_max"]
    fts_locs = data_dict["seg_center"]
    if cfg.dim_loc > 3:                                                         # :117-118,127-132
        query_pos = linear_ln(query_locs[:, :, :3], sd, "coord_encoder.") + linear_ln(query_locs[:, :, 3:6], sd, "box_encoder.")
        fts_pos = linear_ln(fts_locs[:, :, :3], sd, "coord_encoder.") + linear_ln(fts_locs[:, :, 3:6], sd, "box_encoder.")
        fts_pos = fts_pos + linear_ln(fts_locs[:, :, 3:6], sd, "box_encoder.")   # added twice in the reference
    else:
        query_pos = coordinate_encoder(query_locs[:, :, :3], sd, "coord_encoder.", cmin, cmax)
        fts_pos = coordinate_encoder(fts_locs[:, :, :3], sd, "coord_encoder.", cmin, cmax)
    input_dict["query"] = (torch.zeros_like(query_pos), qmask, query_pos)       # :121-123

    def obj_enc(name, x):
        x = linear_ln(x, sd, f"{name}_encoder.input_feat_proj.") if name in cfg.projected_memories else x
        return x if cfg.train is None else cfg.train.obj_dropout(name, x)

    for m in cfg.memories:                                                      # :133-160
        if m == "prompt":
            feat, mask = prompt_encoder(sd, cfg, data_dict)
            pos = None
        elif m in ("mv", "pc"):
            feat = obj_enc(m, data_dict[f"{m}_seg_fts"])
            mask, pos = data_dict[f"{m}_seg_pad_masks"].logical_not(), fts_pos
        elif m == "voxel":
            if "voxel_seg_fts_multiscale" in data_dict:
                feat = list(data_dict["voxel_seg_fts_multiscale"])
                mask = data_dict["seg_pad_masks"].logical_not()
            else:
                feat = obj_enc(m, data_dict["voxel_seg_fts"])
                mask = data_dict["voxel_seg_pad_masks"].logical_not()
            pos = fts_pos
        else:
            raise NotImplementedError(m)
        input_dict[m] = [feat, mask, pos]

    seg_fts_for_match = []                                                      # :167-174
    for m in cfg.memories:
        if m in ("voxel", "mv", "pc"):
            feats = list(input_dict[m])
            if isinstance(feats[0], list):
                feats[0] = feats[0][-1]
            seg_fts_for_match.append(feats)
    seg_masks = data_dict["seg_pad_masks"].logical_not()
    has_mask_head = "mask" in cfg.heads

    def mask_head(query, skip=cfg.skip_query_encoder_mask_pred):
        return mask_head_seg_level(query, sd, "mask_head.", seg_fts_for_match, seg_masks,
                                   cfg.filter_out_classes, None, skip,
                                   None if cfg.train is None else (lambda h: cfg.train.head_dropout("mask", h)))

    pairwise = calc_pairwise_locs(query_locs[:, :, :3]) if cfg.decoder.spatial_selfattn else None  # :182-187
    query, pcls, pmask = query_mask_encoder(sd, cfg.decoder, input_dict, pairwise,
                                            mask_head if has_mask_head else None, prefix="unified_encoder.")
    data_dict["query_feat"] = query
    for head in cfg.heads:                                                      # :193-220
        if head == "mask":
            if cfg.skip_query_encoder_mask_pred:
                pcls, pmask = [], []
            c, m_, _ = mask_head(query, skip=False)
            pcls.append(c)
            pmask.append(m_)
            data_dict["predictions_class"], data_dict["predictions_mask"] = pcls, pmask
        elif head == "ground":
            logits = mlp_head(query, sd, "ground_head.og3d_head.",
                              None if cfg.train is None else (lambda h: cfg.train.head_dropout("ground", h))).squeeze(2)   # grounding_head.py:51-55
            logits = logits.masked_fill(data_dict["query_pad_masks"].logical_not(), float("-inf"))
            data_dict["ground_logits"] = logits
            data_dict["og3d_logits"] = logits
        else:
            raise NotImplementedError(head)
    return data_dict


# =====================================================================================================================
# §8f-2: voxel -> segment pooling (modules/vision/pcd_mask3d_encoder.py:144-154)
# =====================================================================================================================
def scatter_mean(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_scatter.scatter_mean(src, index, dim=0, dim_size=dim_size) restated.  torch_scatter is a third-party
    dependency that is NOT vendored in /root/reference (requirements.txt:74 pins torch_scatter==2.1.2, :79
    torch-scatter==2.1.1; not installed here, no network).  Its published algorithm (torch_scatter/scatter.py,
    `scatter_mean`): out = scatter_sum(src) — `out.scatter_add_(0, index, src)` on a zero tensor —, count =
    scatter_sum(ones).clamp_(min=1), out.true_divide_(count).  On the CPU `scatter_add_` / `index_add_` visit the rows
    in order, so each output row is the fp32 sum of its members in ascending source order.  Anchored on the reference's
    call site pcd_mask3d_encoder.py:150: `self.scatter_fn(f, p2s, dim=0, dim_size=max_seg)`."""
    out = torch.zeros(dim_size, src.shape[1], dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    count = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    count.index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
    return out / count.clamp(min=1).unsqueeze(-1)


def seg_level_pool(feats_per_scene: List[Tensor], point2segment: List[Tensor], max_seg: int, sd: SD, prefix: str) -> Tensor:
    """One scale of PCDMask3DSegLevelEncoder.forward (:146-153): stack(scatter_mean per scene) -> feat_proj =
    Linear + LayerNorm (+ Dropout, identity in eval).  `prefix` = 'feat_proj_list.{i}.'."""
    batch_feat = torch.stack([scatter_mean(f, p, max_seg) for f, p in zip(feats_per_scene, point2segment)])
    return linear_ln(batch_feat, sd, prefix)


# =====================================================================================================================
# §8f-3: matcher cost matrices (modules/third_party/mask3d/matcher.py) and matched mask losses (criterion.py)
# =====================================================================================================================
def batch_dice_cost(inputs: Tensor, targets: Tensor) -> Tensor:
    """matcher.py:12-28."""
    inputs = inputs.sigmoid().flatten(1)
    numerator = 2 * torch.einsum("nc,mc->nm", inputs, targets)
    denominator = inputs.sum(-1)[:, None] + targets.sum(-1)[None, :]
    return 1 - (numerator + 1) / (denominator + 1)


def batch_sigmoid_ce_cost(inputs: Tensor, targets: Tensor) -> Tensor:
    """matcher.py:36-59."""
    hw = inputs.shape[1]
    pos = F.binary_cross_entropy_with_logits(inputs, torch.ones_like(inputs), reduction="none")
    neg = F.binary_cross_entropy_with_logits(inputs, torch.zeros_like(inputs), reduction="none")
    loss = torch.einsum("nc,mc->nm", pos, targets) + torch.einsum("nc,mc->nm", neg, (1 - targets))
    return loss / hw


def matcher_cost(pred_logits: Tensor, pred_masks: Tensor, labels: Tensor, tgt_mask: Tensor, cost_class: float,
                 cost_mask: float, cost_dice: float, ignore_label: int = -100) -> Tensor:
    """One scene of HungarianMatcher.memory_efficient_forward (matcher.py:110-181) with num_points = -1:
    pred_logits (N, C), pred_masks (S, N), labels (M,), tgt_mask (M, S) -> C (N, M)."""
    out_prob = pred_logits.float().softmax(-1)
    tgt_ids = labels.clone()
    filter_ignore = tgt_ids == ignore_label
    tgt_ids[filter_ignore] = 0
    c_class = -out_prob[:, tgt_ids]
    c_class[:, filter_ignore] = -1.0
    out_mask = pred_masks.T.float()
    tgt = tgt_mask.to(out_mask).float()
    return cost_mask * batch_sigmoid_ce_cost(out_mask, tgt) + cost_class * c_class + cost_dice * batch_dice_cost(out_mask, tgt)


def matched_mask_losses(pred_masks: Tensor, tgt_masks: List[Tensor], indices) -> Dict[str, Tensor]:
    """SetCriterion.loss_masks (criterion.py:163-196) with num_points = -1; dice_loss :26-46, sigmoid_ce_loss :51-71.
    pred_masks (B, S, N); tgt_masks[b] (M_b, S); indices[b] = (query idx, target idx)."""
    loss_masks, loss_dices = [], []
    for b, (map_id, target_id) in enumerate(indices):
        mp = pred_masks[b][:, map_id].T
        tm = tgt_masks[b][target_id].float()
        num_masks = tm.shape[0]
        ce = F.binary_cross_entropy_with_logits(mp, tm, reduction="none")
        loss_masks.append(ce.mean(1).sum() / num_masks)
        s = mp.sigmoid().flatten(1)
        numerator = 2 * (s * tm).sum(-1)
        denominator = s.sum(-1) + tm.sum(-1)
        loss_dices.append((1 - (numerator + 1) / (denominator + 1)).sum() / num_masks)
    return {"loss_mask": torch.mean(torch.stack(loss_masks)), "loss_dice": torch.mean(torch.stack(loss_dices))}


# =====================================================================================================================
# §8f-4: inference post-processing (evaluator/instseg_eval.py:85-150, 272-305), one scene, no DBSCAN
# =====================================================================================================================
def instseg_postprocess(pred_logits: Tensor, pred_masks: Tensor, voxel2segment: Tensor, voxel_to_full: Tensor,
                        segment_to_full: Tensor, topk_per_scene: int = -1) -> Dict[str, Tensor]:
    """pred_logits (Q, C+1), pred_masks (S, Q).  Follows eval_instance_step line by line: :97 softmax, :103 segment ->
    voxel gather, :111 drop the no-object class, :121 get_mask_and_scores, :124-131 get_full_res_mask, :137-143 sort."""
    logits = F.softmax(pred_logits, dim=-1)[..., :-1]                      # :89, :111
    masks = pred_masks[voxel2segment]                                      # :103  (V, Q)
    num_queries, num_classes = logits.shape
    labels = torch.arange(num_classes).unsqueeze(0).repeat(num_queries, 1).flatten(0, 1)
    k = num_queries if topk_per_scene == -1 else topk_per_scene
    scores_per_query, topk_indices = logits.flatten(0, 1).topk(k, sorted=True)     # :286-289
    labels_per_query = labels[topk_indices]
    topk_q = torch.div(topk_indices, num_classes, rounding_mode="trunc")
    masks = masks[:, topk_q]
    result_pred_mask = (masks > 0).float()
    heatmap = masks.float().sigmoid()
    mask_scores = (heatmap * result_pred_mask).sum(0) / (result_pred_mask.sum(0) + 1e-6)
    score = scores_per_query * mask_scores

    def full_res(mask, is_heatmap):                                        # :272-281
        mask = mask[voxel_to_full]
        if not is_heatmap:
            mask = scatter_mean(mask, segment_to_full, int(segment_to_full.max()) + 1)
            mask = (mask > 0.5).float()
            mask = mask[segment_to_full]
        return mask
    m_full, h_full = full_res(result_pred_mask, False), full_res(heatmap, True)
    order = score.sort(descending=True)                                    # :137
    return {"scores": order.values, "classes": labels_per_query[order.indices], "masks": m_full[:, order.indices],
            "heatmap": h_full[:, order.indices], "query": topk_q[order.indices]}
