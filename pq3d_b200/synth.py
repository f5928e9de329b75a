"""Seeded synthetic scenes, decoder inputs and weights (the `data_dict` contract of SURVEY.md §8d).

Mirrors what the reference's dataset + collate hand to the model
(data/datasets/sceneverse_instseg.py:199-234, data/datasets/instseg_wrapper.py:59-66): padded
per-scene segment tables, True=valid pad masks, scene bounds, FPS-like query locations.  There is
no network for datasets or checkpoints, so benchmarks and parity tests draw everything from
`torch.Generator().manual_seed(...)` on the CPU — the same image (same torch build) runs here and
on the GPU box, so a (seed, config) pair names identical tensors on both.

Nothing in here touches the oracle or the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch

SCENE_MEMORIES = ("mv", "pc", "voxel")


@dataclass
class Workload:
    """One BASELINE.json config, as shapes."""
    name: str
    B: int                      # scenes (per process)
    N: int                      # queries per scene
    S: int                      # max segment tokens per scene (padded batch max)
    memories: Sequence[str]
    structure: str = "parallel"
    T: int = 0                  # prompt tokens
    num_layers: int = 4
    num_blocks: int = 1
    use_self_mask: bool = False
    spatial_selfattn: bool = True
    ragged: Optional[Sequence[int]] = None     # (lo, hi) for S_b ~ U{lo..hi}
    voxel_multiscale: bool = False
    hidden_size: int = 768
    num_heads: int = 12
    seed: int = 1234

    def decoder_kwargs(self) -> dict:
        return dict(memories=list(self.memories), hidden_size=self.hidden_size,
                    num_attention_heads=self.num_heads, num_layers=self.num_layers,
                    spatial_selfattn=self.spatial_selfattn, structure=self.structure,
                    use_self_mask=self.use_self_mask, num_blocks=self.num_blocks)


def workload(name: str, scenes_per_gpu: Optional[int] = None) -> Workload:
    """BASELINE.json configs[0..4] → per-process shapes (SURVEY.md §8d 'Configs → shapes')."""
    if name == "c1":    # 1 scene, 100 queries, 256 seg tokens, point only (reference CPU case)
        w = Workload("c1", 1, 100, 256, ["pc"], "parallel", seed=1234)
    elif name == "c2":  # batch 8, 100 queries, 1024 seg tokens, voxel+point+image, 1 GPU
        w = Workload("c2", 8, 100, 1024, ["voxel", "mv", "pc"], "parallel", seed=1235)
    elif name == "c3":  # batch 32 over 8 GPUs → 4 scenes/GPU, 2048 seg tokens, all + 32-token prompt
        w = Workload("c3", 4, 100, 2048, ["mv", "pc", "voxel", "prompt"], "mixed", T=32, seed=1236)
    elif name == "c4":  # ragged, 200 queries, masked cross-attention, in-loop mask head
        w = Workload("c4", 4, 200, 4096, ["mv", "pc", "voxel"], "parallel", use_self_mask=True,
                     ragged=(128, 4096), seed=1237)
    elif name == "c5":  # training-step shape: 16 scenes over 8 GPUs
        w = Workload("c5", 2, 100, 2048, ["mv", "pc", "voxel", "prompt"], "mixed", T=32, seed=1238)
    else:
        raise KeyError(name)
    if scenes_per_gpu is not None:
        w.B = scenes_per_gpu
    return w


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def decoder_param_shapes(memories: Sequence[str], hidden_size: int = 768, num_heads: int = 12,
                         num_layers: int = 4, spatial_selfattn: bool = True, structure: str = "parallel",
                         dim_feedforward: int = 2048, **_unused) -> Dict[str, tuple]:
    """The state_dict schema of QueryMaskEncoder, in registration order."""
    D, F_, H = hidden_size, dim_feedforward, num_heads
    out: Dict[str, tuple] = {}
    for i in range(num_layers):
        p = f"unified_encoder.{i}."
        if spatial_selfattn:
            for n in ("w_qs", "w_ks", "w_vs", "fc"):
                out[f"{p}self_attn.self_attn.{n}.weight"] = (D, D)
                out[f"{p}self_attn.self_attn.{n}.bias"] = (D,)
            out[f"{p}self_attn.self_attn.pairwise_loc_fc.weight"] = (H, 5)
            out[f"{p}self_attn.self_attn.pairwise_loc_fc.bias"] = (H,)
        else:
            out[f"{p}self_attn.self_attn.in_proj_weight"] = (3 * D, D)
            out[f"{p}self_attn.self_attn.in_proj_bias"] = (3 * D,)
            out[f"{p}self_attn.self_attn.out_proj.weight"] = (D, D)
            out[f"{p}self_attn.self_attn.out_proj.bias"] = (D,)
        out[f"{p}self_attn.norm.weight"] = (D,)
        out[f"{p}self_attn.norm.bias"] = (D,)
        for j in range(len(memories)):
            q = f"{p}cross_attn_list.{j}."
            out[q + "multihead_attn.in_proj_weight"] = (3 * D, D)
            out[q + "multihead_attn.in_proj_bias"] = (3 * D,)
            out[q + "multihead_attn.out_proj.weight"] = (D, D)
            out[q + "multihead_attn.out_proj.bias"] = (D,)
            out[q + "norm.weight"] = (D,)
            out[q + "norm.bias"] = (D,)
        out[p + "ffn.linear1.weight"] = (F_, D)
        out[p + "ffn.linear1.bias"] = (F_,)
        out[p + "ffn.linear2.weight"] = (D, F_)
        out[p + "ffn.linear2.bias"] = (D,)
        out[p + "ffn.norm.weight"] = (D,)
        out[p + "ffn.norm.bias"] = (D,)
        if structure == "gate":
            out[p + "gate_proj.weight"] = (D, D)
            out[p + "gate_proj.bias"] = (D,)
    return out


def mask_head_param_shapes(n_match: int, hidden_size: int = 768, num_targets: int = 201) -> Dict[str, tuple]:
    """State_dict schema of MaskHeadSegLevel (modules/heads/mask_head.py:12-22)."""
    D = hidden_size
    out = {"cls_head.0.weight": (D, D), "cls_head.0.bias": (D,),
           "cls_head.2.weight": (D,), "cls_head.2.bias": (D,),
           "cls_head.4.weight": (num_targets, D), "cls_head.4.bias": (num_targets,)}
    for j in range(n_match):
        out[f"mask_pred_list.{j}.q_proj.weight"] = (D, D)
        out[f"mask_pred_list.{j}.q_proj.bias"] = (D,)
        out[f"mask_pred_list.{j}.k_proj.weight"] = (D, D)
    return out


def draw_state_dict(shapes: Dict[str, tuple], seed: int, sharp: float = 1.0) -> Dict[str, torch.Tensor]:
    """Reference-like init (MHA in_proj xavier-uniform, every nn.Linear
    N(0, 0.02)), but with biases and LayerNorm affines re-drawn so those paths are exercised, and
    `sharp` scaling the attention in-projections to make the softmax peaky for stress tests."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("in_proj_weight"):
            bound = math.sqrt(6.0 / (shp[0] + shp[1])) * sharp
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif "pairwise_loc_fc" in k:
            t = torch.randn(shp, generator=g) * 0.5
        elif len(shp) == 1 and k.endswith("weight"):        # every 1-D weight is a LayerNorm gain
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1 and k.endswith("bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("gauss_B"):
            t = torch.randn(shp, generator=g)
        elif len(shp) == 2:
            s = 0.02 * (sharp if any(n in k for n in ("w_qs", "w_ks")) else 1.0)
            t = torch.randn(shp, generator=g) * s
        else:
            raise ValueError(f"no init rule for {k} {shp}")
        sd[k] = t
    return sd


def decoder_state_dict(w: Workload, seed: int = 0, sharp: float = 1.0) -> Dict[str, torch.Tensor]:
    return draw_state_dict(decoder_param_shapes(**w.decoder_kwargs()), seed, sharp)


# --------------------------------------------------------------------------------------------
# scenes
# --------------------------------------------------------------------------------------------
def scene_lengths(w: Workload, g: torch.Generator) -> List[int]:
    if w.ragged is None:
        return [w.S] * w.B
    lo, hi = w.ragged
    lens = torch.randint(lo, hi + 1, (w.B,), generator=g).tolist()
    lens[0] = hi                    # batch max is the padded width
    return lens


def make_data_dict(w: Workload, rank: int = 0, voxel_scales: int = 1) -> dict:
    """A padded batch as the collate would produce it; every pad mask True = valid."""
    g = torch.Generator().manual_seed(w.seed + 1000 * rank)
    B, N, D = w.B, w.N, w.hidden_size
    lens = scene_lengths(w, g)
    S = max(lens)
    ar = torch.arange(S)
    seg_valid = ar[None, :] < torch.tensor(lens)[:, None]
    seg_center = torch.rand(B, S, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0])
    coord_min = seg_center.masked_fill(~seg_valid[..., None], float("inf")).amin(1) - 0.05
    coord_max = seg_center.masked_fill(~seg_valid[..., None], float("-inf")).amax(1) + 0.05
    q_locs = torch.zeros(B, N, 3)
    q_valid = torch.ones(B, N, dtype=torch.bool)
    for b in range(B):
        perm = torch.randperm(lens[b], generator=g)
        n = min(N, lens[b])
        q_locs[b, :n] = seg_center[b, perm[:n]]
        q_valid[b, n:] = False
    d = {
        "query_locs": q_locs, "query_pad_masks": q_valid,
        "seg_center": seg_center, "seg_pad_masks": seg_valid,
        "coord_min": coord_min, "coord_max": coord_max,
    }
    for m in w.memories:
        if m == "prompt":
            T = w.T
            tl = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
            d["prompt_feat"] = torch.randn(B, T, D, generator=g)
            d["prompt_pad_masks"] = torch.arange(T)[None, :] < tl[:, None]
        else:
            valid = seg_valid.clone()
            if m == "mv":       # non-suffix validity holes (sceneverse_instseg.py:227)
                valid &= torch.rand(B, S, generator=g) >= 0.10
            d[f"{m}_seg_pad_masks"] = valid
            if m == "voxel" and (w.voxel_multiscale or voxel_scales > 1):
                k = max(voxel_scales, w.num_layers + 1)
                d["voxel_seg_fts_multiscale"] = [torch.randn(B, S, D, generator=g) for _ in range(k)]
            else:
                d[f"{m}_seg_fts"] = torch.randn(B, S, D, generator=g)
    return d


def make_decoder_inputs(w: Workload, rank: int = 0, device="cpu"):
    """`input_dict`/`pairwise_locs` at the QueryMaskEncoder boundary, with stand-in positional
    features (seeded N(0,1) — the coordinate encoder is a producer, not part of the decoder):
    input_dict[name] = [feat, mask(True=ignore), pos], exactly what
    model/query3d_unified.py:110-160 assembles."""
    d = make_data_dict(w, rank)
    g = torch.Generator().manual_seed(w.seed + 1000 * rank + 7)
    B, N, D = w.B, w.N, w.hidden_size
    S = d["seg_center"].shape[1]
    query_pos = torch.randn(B, N, D, generator=g)
    fts_pos = torch.randn(B, S, D, generator=g)
    input_dict = {"query": (torch.zeros(B, N, D), ~d["query_pad_masks"], query_pos)}
    for m in w.memories:
        if m == "prompt":
            input_dict[m] = [d["prompt_feat"], ~d["prompt_pad_masks"], None]
        elif m == "voxel" and "voxel_seg_fts_multiscale" in d:
            input_dict[m] = [list(d["voxel_seg_fts_multiscale"]), ~d["seg_pad_masks"], fts_pos]
        else:
            input_dict[m] = [d[f"{m}_seg_fts"], ~d[f"{m}_seg_pad_masks"], fts_pos]
    pairwise = pairwise_locs_cpu(d["query_locs"]) if w.spatial_selfattn else None

    def mv(x):
        if isinstance(x, torch.Tensor):
            return x.to(device)
        if isinstance(x, (list, tuple)):
            return type(x)(mv(y) for y in x)
        return x
    input_dict = {k: (mv(tuple(v)) if k == "query" else list(mv(list(v)))) for k, v in input_dict.items()}
    return input_dict, (None if pairwise is None else pairwise.to(device)), d


def pairwise_locs_cpu(centers: torch.Tensor, eps: float = 1e-10) -> torch.Tensor:
    """5-D pairwise geometry (modules/utils.py:38-68, 'center', spatial_dist_norm, spatial_dim=5) —
    host-side producer used to build synthetic decoder inputs."""
    d = centers[:, :, None, :] - centers[:, None, :, :]
    dist = torch.sqrt((d ** 2).sum(3) + eps)
    norm = dist / dist.flatten(1).max(dim=1)[0][:, None, None]
    d2 = torch.sqrt((d[..., :2] ** 2).sum(3) + eps)
    return torch.stack([norm, d[..., 2] / dist, d2 / dist, d[..., 1] / d2, d[..., 0] / d2], dim=3)


def clone_input_dict(input_dict: dict) -> dict:
    """The decoder mutates input_dict[m][1] / ['voxel'][0] in place (query_encoder.py:88,91);
    callers that reuse inputs pass a shallow structural copy."""
    out = {}
    for k, v in input_dict.items():
        out[k] = tuple(v) if isinstance(v, tuple) else [list(x) if isinstance(x, list) else x for x in v]
    return out


# --------------------------------------------------------------------------------------------
# whole-model (Query3DUnified) schema
# --------------------------------------------------------------------------------------------
def model_cfg_dict(w: Workload, dim_loc: int = 3, heads: Sequence[str] = ("mask",), voxel_in: int = 128,
                   num_targets: int = 201, filter_out_classes: Sequence[int] = (0, 2),
                   skip_query_encoder_mask_pred: bool = False) -> dict:
    """A plain-dict `cfg` carrying exactly the `cfg.model.*` keys Query3DUnified reads
    (model/query3d_unified.py:31-78; layout of configs/instseg_sceneverse.yaml:93-156)."""
    D = w.hidden_size
    model = {
        "name": "Query3DUnified", "memories": list(w.memories), "hidden_size": D,
        "use_offline_voxel_fts": True, "use_offline_attn_mask": False,
        "skip_query_encoder_mask_pred": skip_query_encoder_mask_pred,
        "obj_loc": {"spatial_dim": 5, "dim_loc": dim_loc, "pairwise_rel_type": "center"},
        "unified_encoder": {"name": "QueryMaskEncoder", "args": w.decoder_kwargs()},
        "heads": list(heads),
    }
    for m in w.memories:
        if m == "prompt":
            continue
        model[f"{m}_encoder"] = {"name": "ObjectEncoder", "args": {
            "input_feat_size": voxel_in if m == "voxel" else D, "hidden_size": D,
            "use_projection": True, "use_cls_head": False, "dropout": 0.1}}
    if "mask" in heads:
        model["mask_head"] = {"name": "MaskHeadSegLevel", "args": {
            "hidden_size": D, "num_targets": num_targets, "memories_for_match": list(w.memories),
            "filter_out_classes": list(filter_out_classes)}}
    if "ground" in heads:
        model["ground_head"] = {"name": "GroundHead", "args": {"hidden_size": 384, "input_size": D, "dropout": 0.3}}
    return {"model": model, "solver": {"lr": 1e-4}}


def model_param_shapes(cfg: dict) -> Dict[str, tuple]:
    m = cfg["model"]
    D = m["hidden_size"]
    out: Dict[str, tuple] = {}
    for mem in m["memories"]:
        if mem == "prompt":
            continue
        a = m[f"{mem}_encoder"]["args"]
        out[f"{mem}_encoder.input_feat_proj.0.weight"] = (D, a["input_feat_size"])
        out[f"{mem}_encoder.input_feat_proj.0.bias"] = (D,)
        out[f"{mem}_encoder.input_feat_proj.1.weight"] = (D,)
        out[f"{mem}_encoder.input_feat_proj.1.bias"] = (D,)
    if m["obj_loc"]["dim_loc"] > 3:
        for e in ("coord_encoder", "box_encoder"):
            out[f"{e}.0.weight"] = (D, 3)
            out[f"{e}.0.bias"] = (D,)
            out[f"{e}.1.weight"] = (D,)
            out[f"{e}.1.bias"] = (D,)
    else:
        out["coord_encoder.pos_enc.gauss_B"] = (3, D // 2)
        out["coord_encoder.feat_proj.0.weight"] = (D, D)
        out["coord_encoder.feat_proj.0.bias"] = (D,)
        out["coord_encoder.feat_proj.1.weight"] = (D,)
        out["coord_encoder.feat_proj.1.bias"] = (D,)
    for k, v in decoder_param_shapes(**m["unified_encoder"]["args"]).items():
        out["unified_encoder." + k] = v
    if "mask" in m["heads"]:
        a = m["mask_head"]["args"]
        n_match = len([x for x in a["memories_for_match"] if x in SCENE_MEMORIES])
        for k, v in mask_head_param_shapes(n_match, D, a["num_targets"]).items():
            out["mask_head." + k] = v
    if "ground" in m["heads"]:
        a = m["ground_head"]["args"]
        hs = a["hidden_size"]
        out.update({"ground_head.og3d_head.0.weight": (hs, D), "ground_head.og3d_head.0.bias": (hs,),
                    "ground_head.og3d_head.2.weight": (hs,), "ground_head.og3d_head.2.bias": (hs,),
                    "ground_head.og3d_head.4.weight": (1, hs), "ground_head.og3d_head.4.bias": (1,)})
    return out


def make_model_data_dict(w: Workload, cfg: dict, rank: int = 0) -> dict:
    """data_dict for Query3DUnified.forward: raw (un-projected) per-modality segment features."""
    d = make_data_dict(w, rank)
    g = torch.Generator().manual_seed(w.seed + 1000 * rank + 13)
    m = cfg["model"]
    B, S = d["seg_center"].shape[:2]
    if "voxel" in w.memories:
        cin = m["voxel_encoder"]["args"]["input_feat_size"]
        d["voxel_seg_fts"] = torch.randn(B, S, cin, generator=g)
        d.pop("voxel_seg_fts_multiscale", None)
    if m["obj_loc"]["dim_loc"] > 3:
        d["query_locs"] = torch.cat([d["query_locs"], torch.rand(B, w.N, 3, generator=g)], dim=2)
        d["seg_center"] = torch.cat([d["seg_center"], torch.rand(B, S, 3, generator=g)], dim=2)
    return d
