"""CPU, this container only (skipped where /root/reference is absent): the drop-in boundary against the LIVE reference.

  * SURVEY §8a row 12 — init parity: `_reference_init` reproduces what modules/utils.py:28-32 (layer_repeat deep
    copies), the sub-layers' `_reset_parameters` (xavier on dim > 1) and modules/weights.py:3-19 (_init_weights_bert)
    leave behind: identical xavier-uniform `in_proj_weight` in every copy, N(0, 0.02) Linears with zero bias,
    LayerNorm (1, 0), zero `in_proj_bias`.
  * SURVEY §8b — the registry plugin: `pq3d_b200.registry.register_into_reference()` swaps the hot-path classes
    into the reference's own registries; the model is then built by the reference's `build_model` /
    `build_module_by_name` (model/build.py:17-19, modules/build.py:24-31) and must expose the reference's
    `state_dict` schema and `get_opt_params()` grouping (model/query3d_unified.py:224-238, optim/utils.py:1-18).
"""
import math

import pytest
import torch

from oracle import ref_loader
from pq3d_b200 import synth

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


def _stats(sd):
    """Per-kind summary of an initialised decoder state_dict."""
    out = {}
    for k, v in sd.items():
        if k.endswith("in_proj_weight"):
            kind = "in_proj_weight"
        elif k.endswith("in_proj_bias"):
            kind = "in_proj_bias"
        elif ".norm." in k:
            kind = "norm." + k.rsplit(".", 1)[1]
        elif k.endswith(".bias"):
            kind = "linear.bias"
        else:
            kind = "linear.weight"
        out.setdefault(kind, []).append((k, v))
    return out


@needs_ref
@pytest.mark.parametrize("kw", [
    dict(memories=["mv", "pc", "voxel", "prompt"], spatial_selfattn=True, structure="mixed", num_layers=3),
    dict(memories=["pc", "voxel"], spatial_selfattn=False, structure="parallel", num_layers=2),
])
def test_init_parity_with_live_reference(kw):
    from pq3d_b200.query_encoder import QueryMaskEncoder
    ns = ref_loader.load()
    torch.manual_seed(7)
    ref = ns.query_encoder.QueryMaskEncoder(None, **kw)
    torch.manual_seed(8)
    ours = QueryMaskEncoder(None, **kw)
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr) == list(so) and all(sr[k].shape == so[k].shape for k in sr)
    D = 768
    for name, sd in (("reference", sr), ("ours", so)):
        st = _stats(sd)
        # (1) the bare MHA in_proj_weight keeps xavier-uniform and is IDENTICAL in every deep copy of a sub-layer
        #     (per kind: all cross-attention copies alike, all self-attention copies alike)
        for kind_sel in (".cross_attn_list.", ".self_attn.self_attn.in_proj"):
            ws = [v for k, v in st.get("in_proj_weight", []) if kind_sel in k]
            for w in ws[1:]:
                assert torch.equal(w, ws[0]), f"{name}: in_proj_weight copies differ ({kind_sel})"
            if ws:
                bound = math.sqrt(6.0 / (D + 3 * D))
                assert ws[0].abs().max() <= bound + 1e-7
                assert abs(ws[0].std().item() - bound / math.sqrt(3.0)) <= 0.01 * bound
                assert abs(ws[0].mean().item()) <= 1e-3 * bound * 10
        for k, v in st.get("in_proj_bias", []):
            assert (v == 0).all(), k
        # (2) every nn.Linear re-drawn N(0, 0.02), bias 0 — independent draws, so copies are NOT identical
        lw = [v for k, v in st["linear.weight"]]
        for v in lw:
            if v.numel() >= 768 * 12:
                assert abs(v.std().item() - 0.02) <= 0.002, name
        big = [v for v in lw if v.shape == (D, D)]
        assert not torch.equal(big[0], big[1]), f"{name}: Linear weights must be independent draws"
        for k, v in st["linear.bias"]:
            assert (v == 0).all(), k
        # (3) LayerNorm (1, 0)
        for k, v in st["norm.weight"]:
            assert (v == 1).all(), k
        for k, v in st["norm.bias"]:
            assert (v == 0).all(), k
    # the two implementations agree on the distribution parameters they realise
    a, b = _stats(sr), _stats(so)
    ra = torch.cat([v.flatten() for _, v in a["linear.weight"]]).std().item()
    rb = torch.cat([v.flatten() for _, v in b["linear.weight"]]).std().item()
    assert abs(ra - rb) <= 2e-4
    xa = [v for k, v in a.get("in_proj_weight", []) if ".cross_attn_list." in k][0]
    xb = [v for k, v in b.get("in_proj_weight", []) if ".cross_attn_list." in k][0]
    assert abs(xa.std().item() - xb.std().item()) <= 2e-4 and abs(xa.abs().max().item() - xb.abs().max().item()) <= 1e-4


def _registries(ns):
    from model import build as model_build
    regs = [ns.build.GROUNDING_REGISTRY, ns.build.HEADS_REGISTRY, ns.build.VISION_REGISTRY, model_build.MODEL_REGISTRY]
    return regs, [dict(r._map) for r in regs]


@needs_ref
@pytest.mark.parametrize("stage", ["stage1", "stage2"])
def test_registry_plugin_builds_through_reference_and_matches_schema(stage):
    from pq3d_b200 import registry
    from pq3d_b200.query3d_unified import Query3DUnified as Ours
    from pq3d_b200.query_encoder import QueryMaskEncoder as OursEnc
    ns = ref_loader.load()
    if stage == "stage1":
        w = synth.Workload("p1", 2, 20, 48, ["mv", "pc", "voxel"], "parallel", num_layers=2, use_self_mask=True,
                           num_blocks=2)
        cfg = synth.model_cfg_dict(w, dim_loc=3, heads=("mask",))
    else:
        w = synth.Workload("p2", 2, 20, 48, ["mv", "pc", "voxel", "prompt"], "mixed", T=6, num_layers=2)
        cfg = synth.model_cfg_dict(w, dim_loc=6, heads=("ground",))
        cfg["model"]["ground_head"]["lr"] = 5e-4                      # a per-module lr override (:229-232)

        class _StubTxt(torch.nn.Module):                              # the CLIP tower is out of scope: parameter-free stub
            def __init__(self, cfg=None, **kw):
                super().__init__()
        cfg["model"]["txt_encoder"] = {"name": "_StubTxt"}
        ns.build.LANGUAGE_REGISTRY._map["_StubTxt"] = _StubTxt
    acfg = ref_loader.to_attr(cfg)
    ref_model = ns.query3d_unified.Query3DUnified(acfg)                # built while the registries are still stock
    regs, saved = _registries(ns)
    try:
        done = registry.register_into_reference()
        assert "grounding:QueryMaskEncoder" in done and "model:Query3DUnified" in done
        # the reference's own factory functions now return the B200 classes
        enc = ns.build.build_module_by_name(acfg.model.unified_encoder)
        assert isinstance(enc, OursEnc)
        model = ns.model_build.build_model(acfg)
        assert isinstance(model, Ours) and isinstance(model.unified_encoder, OursEnc)
    finally:
        for r, m in zip(regs, saved):
            r._map.clear()
            r._map.update(m)
    # state_dict schema: same keys, same order, same shapes (a reference checkpoint loads with strict=True)
    sr, so = ref_model.state_dict(), model.state_dict()
    assert list(sr) == list(so)
    assert all(tuple(sr[k].shape) == tuple(so[k].shape) for k in sr)
    model.load_state_dict(sr, strict=True)
    assert enc.state_dict().keys() == ref_model.unified_encoder.state_dict().keys()
    # get_opt_params: same groups (module name, lr, weight decay) holding the same parameter names in the same order
    def named_groups(m):
        names = {id(p): n for n, p in m.named_parameters()}
        return [(g["name"], g["lr"], g["weight_decay"], [names[id(p)] for p in g["params"]]) for g in m.get_opt_params()
                if g["name"] != "txt_encoder"]                         # the text tower is not part of this package
    gr, go = named_groups(ref_model), named_groups(model)
    assert gr == go
    if stage == "stage2":
        assert any(name == "ground_head" and lr == 5e-4 for name, lr, _, _ in go)
    # decoder LayerNorm weights are named `norm.weight`, so they fall into the decay group
    decay = [n for name, _, wd, ns_ in go if wd > 0 for n in ns_]
    assert any(n.endswith("cross_attn_list.0.norm.weight") for n in decay)
