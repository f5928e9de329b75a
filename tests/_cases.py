"""Shared helpers: rebuild a golden case's weights/inputs from its recipe and run the oracle."""
import glob
import os
from functools import partial

import torch

from oracle import restatement as O
from pq3d_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(kind=None):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.pt"))):
        name = os.path.basename(f)[:-3]
        k = "decoder" if name.startswith("dec_") else "maskhead" if name.startswith("mh_") else "model"
        if kind is None or k == kind:
            out.append(name)
    return out


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def build_workload(case) -> synth.Workload:
    w = synth.workload(case["base"])
    for k, v in case["over"].items():
        setattr(w, k, v)
    return w


def to_dev(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, dict):
        return {k: to_dev(v, device) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_dev(v, device) for v in x)
    return x


def mask_head_inputs(w, inp):
    feats = []
    for m in w.memories:
        if m in synth.SCENE_MEMORIES:
            f = list(inp[m])
            if isinstance(f[0], list):
                f[0] = f[0][-1]
            feats.append(f)
    return feats


def oracle_decoder(case, device="cpu"):
    w = build_workload(case)
    sd = to_dev(synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"]), device)
    inp, pw, _ = synth.make_decoder_inputs(w, device=device)
    cfg = O.DecoderCfg(**w.decoder_kwargs())
    with torch.no_grad():
        q, _, _ = O.query_mask_encoder(sd, cfg, inp, pw)
    return {"query": q}


def oracle_maskhead(case, device="cpu"):
    w = build_workload(case)
    sd = to_dev(synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"]), device)
    n_match = len([m for m in w.memories if m in synth.SCENE_MEMORIES])
    sd_mh = to_dev(synth.draw_state_dict(synth.mask_head_param_shapes(n_match), case["wseed"] + 100), device)
    inp, pw, d = synth.make_decoder_inputs(w, device=device)
    head = partial(O.mask_head_seg_level, sd=sd_mh, prefix="", seg_fts_for_match=mask_head_inputs(w, inp),
                   seg_masks=(~d["seg_pad_masks"]).to(device), filter_out_classes=[0, 2])
    cfg = O.DecoderCfg(**w.decoder_kwargs())
    with torch.no_grad():
        q, pc, pm = O.query_mask_encoder(sd, cfg, inp, pw, head)
        c, m, a = head(q)
    return {"query": q, "pred_class_last": pc[-1], "pred_mask_last": pm[-1], "pred_mask_first": pm[0],
            "final_class": c, "final_mask": m, "final_attn_mask": a, "n_pred": torch.tensor(len(pm))}


def model_cfg(case):
    w = build_workload(case)
    cfg = synth.model_cfg_dict(w, dim_loc=case["dim_loc"], heads=case["heads"],
                               skip_query_encoder_mask_pred=case["skip"])
    return w, cfg


def oracle_model_cfg(w, cfg) -> O.ModelCfg:
    m = cfg["model"]
    return O.ModelCfg(memories=list(w.memories), decoder=O.DecoderCfg(**w.decoder_kwargs()),
                      dim_loc=m["obj_loc"]["dim_loc"], heads=tuple(m["heads"]),
                      skip_query_encoder_mask_pred=m["skip_query_encoder_mask_pred"],
                      filter_out_classes=(m["mask_head"]["args"]["filter_out_classes"] if "mask" in m["heads"] else None))


def oracle_model(case, device="cpu"):
    w, cfg = model_cfg(case)
    sd = to_dev(synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"]), device)
    d = to_dev(synth.make_model_data_dict(w, cfg), device)
    with torch.no_grad():
        out = O.query3d_unified_forward(sd, oracle_model_cfg(w, cfg), d)
    res = {}
    if "mask" in case["heads"]:
        res["pred_class_last"] = out["predictions_class"][-1]
        res["pred_mask_last"] = out["predictions_mask"][-1]
        res["n_pred"] = torch.tensor(len(out["predictions_mask"]))
    if "ground" in case["heads"]:
        res["ground_logits"] = out["ground_logits"]
    return res


def assert_close_to_golden(res, gold, rtol=1e-5):
    for k, g in gold.items():
        r = res[k].cpu()
        if g.dtype == torch.bool:
            assert torch.equal(r, g), f"{k}: bool mismatch in {(r != g).sum().item()} places"
        elif g.ndim == 0:
            assert int(r) == int(g), k
        else:
            fin = torch.isfinite(g)
            assert torch.equal(torch.isfinite(r), fin), f"{k}: non-finite pattern differs"
            # -1e6 fill values are exact; compare the finite part relative to its own scale
            scale = g[fin & (g.abs() < 1e5)].abs().max().clamp_min(1e-6)
            err = (r[fin] - g[fin]).abs().max() / scale
            assert err <= rtol, f"{k}: rel err {err:.3e} > {rtol}"
