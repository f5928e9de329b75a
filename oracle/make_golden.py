"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the REAL reference modules.

    python -m oracle.make_golden            # needs /root/reference (this container only)

The reference tree holds no fixtures for the decoder path, so these are outputs of
the reference itself: each case instantiates the reference's own `QueryMaskEncoder` /
`MaskHeadSegLevel` / `Query3DUnified` (imported by oracle/ref_loader.py), loads weights drawn by
`pq3d_b200.synth.draw_state_dict(seed)`, feeds seeded synthetic inputs, and stores only the
OUTPUTS plus the (seed, shape) recipe — weights and inputs are re-drawn from the recipe at test
time, on the same torch build, so the fixtures stay a few hundred kB.
"""
from __future__ import annotations

import os
from functools import partial

import torch
import torch.nn as nn

from oracle import ref_loader
from pq3d_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (workload overrides, weight seed, sharp)
DECODER_CASES = {
    "dec_c1": dict(base="c1", over={}, wseed=0, sharp=1.0),
    "dec_c1_sharp": dict(base="c1", over=dict(N=37, S=130), wseed=1, sharp=4.0),
    "dec_parallel3": dict(base="c2", over=dict(B=2, N=40, S=152, num_layers=2), wseed=2, sharp=2.0),
    "dec_mixed": dict(base="c3", over=dict(B=2, N=50, S=200, T=8, num_layers=2), wseed=3, sharp=2.0),
    "dec_sequential": dict(base="c3", over=dict(B=2, N=33, S=96, T=5, num_layers=1, structure="sequential"), wseed=4, sharp=2.0),
    "dec_gate": dict(base="c3", over=dict(B=1, N=20, S=64, T=6, num_layers=1, structure="gate"), wseed=5, sharp=2.0),
    "dec_plain_selfattn": dict(base="c2", over=dict(B=2, N=24, S=80, num_layers=1, spatial_selfattn=False), wseed=6, sharp=2.0),
    "dec_voxel_multiscale": dict(base="c2", over=dict(B=1, N=30, S=72, num_layers=3, voxel_multiscale=True), wseed=7, sharp=2.0),
}
# decoder + in-loop mask head (stage-1 style: use_self_mask, num_blocks>1)
MASKHEAD_CASES = {
    "mh_selfmask": dict(base="c4", over=dict(B=2, N=40, S=150, ragged=(60, 150), num_layers=2, num_blocks=2,
                                              voxel_multiscale=True), wseed=8, sharp=2.0),
}
MODEL_CASES = {
    "model_stage1": dict(base="c4", over=dict(B=2, N=32, S=120, ragged=(50, 120), num_layers=2, num_blocks=2),
                         wseed=9, sharp=2.0, dim_loc=3, heads=("mask",), skip=False),
    "model_stage2": dict(base="c3", over=dict(B=2, N=24, S=24, T=6, num_layers=2), wseed=10, sharp=2.0,
                         dim_loc=6, heads=("ground",), skip=True),
}


def build_workload(case) -> synth.Workload:
    w = synth.workload(case["base"])
    for k, v in case["over"].items():
        setattr(w, k, v)
    return w


def run_decoder_case(ns, case):
    w = build_workload(case)
    sd = synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"])
    enc = ns.query_encoder.QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
    enc.load_state_dict(sd, strict=True)
    inp, pw, _ = synth.make_decoder_inputs(w)
    with torch.no_grad():
        q, _, _ = enc(inp, pw)
    return {"query": q.contiguous()}


def mask_head_inputs(w, inp):
    feats = []
    for m in w.memories:
        if m in synth.SCENE_MEMORIES:
            f = list(inp[m])
            if isinstance(f[0], list):
                f[0] = f[0][-1]
            feats.append(f)
    return feats


def run_maskhead_case(ns, case):
    w = build_workload(case)
    sd = synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"])
    n_match = len([m for m in w.memories if m in synth.SCENE_MEMORIES])
    sd_mh = synth.draw_state_dict(synth.mask_head_param_shapes(n_match), case["wseed"] + 100)
    enc = ns.query_encoder.QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
    enc.load_state_dict(sd, strict=True)
    mh = ns.mask_head.MaskHeadSegLevel(None, 768, 201, memories_for_match=list(w.memories), filter_out_classes=[0, 2]).eval()
    mh.load_state_dict(sd_mh, strict=True)
    inp, pw, d = synth.make_decoder_inputs(w)
    head = partial(mh, seg_fts_for_match=mask_head_inputs(w, inp), seg_masks=~d["seg_pad_masks"],
                   offline_attn_masks=None, skip_prediction=False)
    with torch.no_grad():
        q, pc, pm = enc(inp, pw, head)
        c, m, a = head(query=q)
    return {"query": q, "pred_class_last": pc[-1], "pred_mask_last": pm[-1], "pred_mask_first": pm[0],
            "final_class": c, "final_mask": m, "final_attn_mask": a, "n_pred": torch.tensor(len(pm))}


class _NullTxt(nn.Module):
    def __init__(self, cfg, **kw):
        super().__init__()


def run_model_case(ns, case):
    w = build_workload(case)
    cfg = synth.model_cfg_dict(w, dim_loc=case["dim_loc"], heads=case["heads"],
                               skip_query_encoder_mask_pred=case["skip"])
    if "prompt" in w.memories:
        cfg["model"]["txt_encoder"] = {"name": "_NullTxt"}
        ns.build.LANGUAGE_REGISTRY._map["_NullTxt"] = _NullTxt
    sd = synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"])
    model = ns.query3d_unified.Query3DUnified(ref_loader.to_attr(cfg)).eval()
    missing = model.load_state_dict(sd, strict=True)
    # CLIP text tower is out of scope: prompt features enter as an opaque (B,T,768) tensor
    model.prompt_encoder = lambda dd: (dd["prompt_feat"], dd["prompt_pad_masks"].logical_not())
    d = synth.make_model_data_dict(w, cfg)
    if "ground" in case["heads"]:
        d["tgt_object_id"] = torch.zeros(w.B, dtype=torch.long)
    with torch.no_grad():
        out = model(d)
    res = {}
    if "mask" in case["heads"]:
        res["pred_class_last"] = out["predictions_class"][-1]
        res["pred_mask_last"] = out["predictions_mask"][-1]
        res["n_pred"] = torch.tensor(len(out["predictions_mask"]))
    if "ground" in case["heads"]:
        res["ground_logits"] = out["ground_logits"]
    return res


def main():
    ns = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for name, case in DECODER_CASES.items():
        res = run_decoder_case(ns, case)
        torch.save({"case": case, "kind": "decoder", "out": res}, os.path.join(OUT, name + ".pt"))
        print(name, {k: tuple(v.shape) for k, v in res.items()})
    for name, case in MASKHEAD_CASES.items():
        res = run_maskhead_case(ns, case)
        torch.save({"case": case, "kind": "maskhead", "out": res}, os.path.join(OUT, name + ".pt"))
        print(name, {k: tuple(v.shape) for k, v in res.items()})
    for name, case in MODEL_CASES.items():
        res = run_model_case(ns, case)
        torch.save({"case": case, "kind": "model", "out": res}, os.path.join(OUT, name + ".pt"))
        print(name, {k: tuple(v.shape) for k, v in res.items()})


if __name__ == "__main__":
    main()
