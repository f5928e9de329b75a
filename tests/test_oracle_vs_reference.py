"""CPU, this container only: the restatement equals the REAL reference modules imported from
/root/reference (skipped on the GPU box, where the tree is absent), plus the known-answer
properties of SURVEY.md §8c that need no reference at all."""
import pytest
import torch

import _cases as C
from oracle import ref_loader, restatement as O
from pq3d_b200 import synth

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("base,over", [
    ("c1", dict(N=64, S=96, num_layers=2)),
    ("c3", dict(B=2, N=30, S=70, T=7, num_layers=2)),
    ("c2", dict(B=2, N=17, S=40, num_layers=1, structure="sequential")),
])
def test_decoder_matches_live_reference(base, over):
    ns = ref_loader.load()
    case = dict(base=base, over=over, wseed=21, sharp=3.0)
    w = C.build_workload(case)
    sd = synth.decoder_state_dict(w, seed=21, sharp=3.0)
    enc = ns.query_encoder.QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
    enc.load_state_dict(sd, strict=True)
    inp, pw, _ = synth.make_decoder_inputs(w)
    with torch.no_grad():
        ref, _, _ = enc(synth.clone_input_dict(inp), pw)
    out = C.oracle_decoder(case)["query"]
    assert (ref - out).abs().max() / ref.abs().max() <= 1e-5


@needs_ref
def test_pairwise_locs_matches_live_reference():
    ns = ref_loader.load()
    c = torch.rand(3, 50, 3, generator=torch.Generator().manual_seed(3)) * 5
    ref = ns.utils.calc_pairwise_locs(c, None, pairwise_rel_type="center", spatial_dist_norm=True, spatial_dim=5)
    assert torch.allclose(O.calc_pairwise_locs(c), ref, rtol=1e-6, atol=1e-7)
    assert torch.allclose(synth.pairwise_locs_cpu(c), ref, rtol=1e-6, atol=1e-7)


@needs_ref
def test_state_dict_schema_matches_live_reference():
    ns = ref_loader.load()
    for kw in (dict(memories=["mv", "pc", "voxel", "prompt"], spatial_selfattn=True, structure="mixed"),
               dict(memories=["pc"], spatial_selfattn=False, structure="gate", num_layers=2)):
        enc = ns.query_encoder.QueryMaskEncoder(None, **kw)
        ref = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
        assert synth.decoder_param_shapes(**kw) == ref
        assert list(synth.decoder_param_shapes(**kw)) == list(ref)          # same order too


class TorchRngTrain:
    """cfg.train hook that draws every random decision from torch's global RNG in the reference's own order and
    shapes: F.dropout per call site, torch.rand for the memory-dropout mask."""

    def __init__(self, p, p_mem, spatial):
        self.p, self.p_mem, self.spatial = p, p_mem, spatial

    def sublayer(self, layer, kind, x):
        if isinstance(kind, tuple) or (kind == "sa" and not self.spatial):
            # nn.MultiheadAttention(batch_first=True) returns a transposed view of an (L, B, E) buffer and nn.Dropout
            # fills its noise in memory order: reproduce that layout so the same draws land on the same elements
            x = x.transpose(0, 1).contiguous().transpose(0, 1)
        return torch.nn.functional.dropout(x, self.p, True)

    def probs(self, layer, kind, P):
        return torch.nn.functional.dropout(P, self.p, True)

    def hidden(self, layer, h):
        return torch.nn.functional.dropout(h, self.p, True)

    def memory_keep(self, layer, memories, B):
        return torch.rand(B, len(memories)) > self.p_mem


@needs_ref
@pytest.mark.parametrize("structure,spatial,p_mem", [("mixed", True, 0.6), ("parallel", False, 0.5),
                                                     ("sequential", True, 0.0)])
def test_training_mode_matches_live_reference(structure, spatial, p_mem):
    """module.train(): sublayer / attention-probability / FFN dropout (p = 0.1) and memory dropout.  With the same
    torch seed the oracle's training branch consumes the RNG exactly like the reference modules do, so outputs AND
    gradients agree to fp32 rounding."""
    ns = ref_loader.load()
    mems = ["mv", "pc", "voxel"] + (["prompt"] if structure != "parallel" else [])
    w = synth.Workload("t", 2, 13, 37, mems, structure, T=5, num_layers=2, spatial_selfattn=spatial)
    kw = dict(w.decoder_kwargs(), memory_dropout=p_mem)
    sd = synth.decoder_state_dict(w, seed=31, sharp=2.0)
    enc = ns.query_encoder.QueryMaskEncoder(None, **kw).train()
    enc.load_state_dict(sd, strict=True)
    inp, pw, _ = synth.make_decoder_inputs(w)
    torch.manual_seed(77)
    ref, _, _ = enc(synth.clone_input_dict(inp), pw)
    ref.square().sum().backward()
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    cfg = O.DecoderCfg(**kw)
    cfg.train = TorchRngTrain(0.1, p_mem, spatial)
    torch.manual_seed(77)
    out, _, _ = O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw)
    out.square().sum().backward()
    assert (ref - out).abs().max() / ref.abs().max() <= 1e-5
    gmax = max(float(p_ref.grad.abs().max()) for _, p_ref in enc.named_parameters())
    for k, p_ref in enc.named_parameters():
        g = sdd[k].grad       # key biases have an exactly-zero true gradient (softmax shift invariance): global floor
        assert g is not None and float((g - p_ref.grad).abs().max()) <= 1e-4 * max(float(p_ref.grad.abs().max()), 1e-3 * gmax), k


# ---- known-answer properties ---------------------------------------------------------------
def _one_ca(S=40, N=9, seed=5):
    w = synth.Workload("t", 2, N, S, ["pc"], "parallel", num_layers=1)
    sd = synth.decoder_state_dict(w, seed=seed, sharp=3.0)
    inp, pw, _ = synth.make_decoder_inputs(w)
    return w, sd, inp


def test_all_keys_masked_returns_ln_of_bias():
    """add_zero_attn: with every key masked the only attended key is the zero one, so the update is
    out_proj.bias and the layer returns LN(tgt + out_proj.bias)."""
    w, sd, inp = _one_ca()
    p = "unified_encoder.0.cross_attn_list.0."
    tgt = torch.randn(2, w.N, 768, generator=torch.Generator().manual_seed(1))
    feat, mask, pos = inp["pc"]
    out = O.cross_attention_layer(tgt, feat, sd, p, 12, None, torch.ones_like(mask), pos, inp["query"][2])
    exp = O.layer_norm(tgt + sd[p + "multihead_attn.out_proj.bias"], sd, p + "norm.")
    assert torch.allclose(out, exp, atol=1e-5)


def test_padding_invariance_and_permutation_equivariance():
    w, sd, inp = _one_ca()
    p = "unified_encoder.0.cross_attn_list.0."
    tgt = torch.randn(2, w.N, 768, generator=torch.Generator().manual_seed(2))
    feat, mask, pos = inp["pc"]
    qp = inp["query"][2]
    base = O.cross_attention_layer(tgt, feat, sd, p, 12, None, mask, pos, qp)
    # appending masked tokens changes nothing
    feat2 = torch.cat([feat, torch.randn(2, 7, 768)], 1)
    pos2 = torch.cat([pos, torch.randn(2, 7, 768)], 1)
    mask2 = torch.cat([mask, torch.ones(2, 7, dtype=torch.bool)], 1)
    assert torch.allclose(O.cross_attention_layer(tgt, feat2, sd, p, 12, None, mask2, pos2, qp), base, atol=2e-6)
    # permuting segment order (with masks alike) changes nothing
    perm = torch.randperm(feat.shape[1], generator=torch.Generator().manual_seed(3))
    out = O.cross_attention_layer(tgt, feat[:, perm], sd, p, 12, None, mask[:, perm], pos[:, perm], qp)
    assert torch.allclose(out, base, atol=2e-6)


def test_parallel_with_one_memory_equals_sequential():
    for structure in ("parallel", "sequential"):
        pass
    wp = synth.Workload("t", 2, 11, 30, ["pc"], "parallel", num_layers=2)
    ws = synth.Workload("t", 2, 11, 30, ["pc"], "sequential", num_layers=2)
    sd = synth.decoder_state_dict(wp, seed=9)
    inp, pw, _ = synth.make_decoder_inputs(wp)
    a, _, _ = O.query_mask_encoder(sd, O.DecoderCfg(**wp.decoder_kwargs()), synth.clone_input_dict(inp), pw)
    b, _, _ = O.query_mask_encoder(sd, O.DecoderCfg(**ws.decoder_kwargs()), synth.clone_input_dict(inp), pw)
    assert torch.allclose(a, b, atol=1e-6)


def test_fully_masked_rows_become_fully_visible():
    """QueryMaskEncoder: attn_mask[attn_mask.all(-1)] = False (query_encoder.py:83)."""
    w = synth.Workload("t", 1, 6, 20, ["pc"], "parallel", num_layers=1, use_self_mask=True)
    sd = synth.decoder_state_dict(w, seed=4)
    inp, pw, _ = synth.make_decoder_inputs(w)
    am = torch.zeros(1, 6, 20, dtype=torch.bool)
    am[0, 2] = True                                   # row 2 would see nothing
    head_all = lambda q: (None, None, am.clone())
    am2 = am.clone(); am2[0, 2] = False
    head_vis = lambda q: (None, None, am2.clone())
    a, _, _ = O.query_mask_encoder(sd, O.DecoderCfg(**w.decoder_kwargs()), synth.clone_input_dict(inp), pw, head_all)
    b, _, _ = O.query_mask_encoder(sd, O.DecoderCfg(**w.decoder_kwargs()), synth.clone_input_dict(inp), pw, head_vis)
    assert torch.equal(a, b)


@needs_ref
@pytest.mark.parametrize("dim_loc", [3, 6])
def test_prompt_encoder_loc_and_txt_rows_match_live_reference(dim_loc):
    """Query3DUnified.prompt_encoder on a mixed batch: location prompts through the coordinate (+ box) encoder,
    broadcast over the prompt slots with only slot 0 valid; text rows from a stub tower that returns fixed features
    (CLIP is out of scope).  The oracle must reproduce features, masks and the in-place mask write-back."""
    import torch.nn as nn
    from oracle import make_golden
    ns = ref_loader.load()
    case = dict(base="c3", over=dict(B=4, N=12, S=20, T=6, num_layers=1), dim_loc=dim_loc, heads=["ground"], skip=False,
                wseed=3, sharp=1.0)
    w, cfg = C.model_cfg(case)
    g = torch.Generator().manual_seed(12)
    txt_feat = torch.randn(w.B, w.T, 768, generator=g)

    class _StubTxt(nn.Module):
        def __init__(self, cfg=None, **kw):
            super().__init__()
            self.rows = None

        def forward(self, ids, mask):
            return txt_feat[self.rows]
    cfg["model"]["txt_encoder"] = {"name": "_StubTxt"}
    ns.build.LANGUAGE_REGISTRY._map["_StubTxt"] = _StubTxt
    sd = synth.draw_state_dict(synth.model_param_shapes(cfg), 3, 1.0)
    model = ns.query3d_unified.Query3DUnified(ref_loader.to_attr(cfg)).eval()
    model.load_state_dict(sd, strict=True)
    ptype = torch.tensor([1, 3, 3, 1])
    model.txt_encoder.rows = ptype == 1
    prompt = torch.rand(w.B, w.T, generator=g) * 3
    pad = torch.rand(w.B, w.T, generator=g) < 0.7
    pad[:, 0] = True
    d = dict(prompt=prompt.clone(), prompt_pad_masks=pad.clone(), prompt_type=ptype,
             coord_min=torch.zeros(w.B, 3), coord_max=torch.full((w.B, 3), 4.0))
    with torch.no_grad():
        ref_feat, ref_mask = model.prompt_encoder(d)
    d2 = dict(prompt=prompt.clone(), prompt_pad_masks=pad.clone(), prompt_type=ptype, prompt_feat=txt_feat,
              coord_min=torch.zeros(w.B, 3), coord_max=torch.full((w.B, 3), 4.0))
    with torch.no_grad():
        feat, mask = O.prompt_encoder(sd, C.oracle_model_cfg(w, cfg), d2)
    assert torch.equal(mask, ref_mask) and torch.equal(d2["prompt_pad_masks"], d["prompt_pad_masks"])
    assert torch.allclose(feat, ref_feat, atol=1e-5)
    assert torch.equal(feat[1, 0], feat[1, w.T - 1])                  # broadcast over the slots


@needs_ref
@pytest.mark.parametrize("stage", ["stage1_mask", "stage2_ground"])
def test_model_training_mode_matches_live_reference(stage):
    """Query3DUnified in .train() — ObjectEncoder dropout, the decoder's dropouts, the mask head's / ground head's MLP
    dropout — against the oracle's model forward whose hooks replay torch's RNG in the reference's own order: outputs
    and the gradient of every parameter agree to fp32 rounding.  (Text tower stubbed: CLIP is out of scope.)"""
    import torch.nn as nn
    ns = ref_loader.load()
    if stage == "stage1_mask":
        case = dict(base="c2", over=dict(B=2, N=14, S=40, num_layers=2, use_self_mask=True), dim_loc=3, heads=["mask"],
                    skip=False, wseed=51, sharp=1.0)
    else:
        case = dict(base="c3", over=dict(B=2, N=14, S=40, T=5, num_layers=2), dim_loc=6, heads=["ground"], skip=False,
                    wseed=52, sharp=1.0)
    w, cfg = C.model_cfg(case)
    g = torch.Generator().manual_seed(2)
    txt_feat = torch.randn(w.B, max(w.T, 1), 768, generator=g)

    class _StubTxt(nn.Module):
        def __init__(self, cfg=None, **kw):
            super().__init__()
    if "prompt" in w.memories:
        cfg["model"]["txt_encoder"] = {"name": "_StubTxt2"}
        ns.build.LANGUAGE_REGISTRY._map["_StubTxt2"] = _StubTxt
    sd = synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"])
    model = ns.query3d_unified.Query3DUnified(ref_loader.to_attr(cfg)).train()
    model.load_state_dict(sd, strict=True)
    model.prompt_encoder = lambda dd: (dd["prompt_feat"], dd["prompt_pad_masks"].logical_not())
    d = synth.make_model_data_dict(w, cfg)
    if "ground" in case["heads"]:
        d["tgt_object_id"] = torch.zeros(w.B, dtype=torch.long)
    clone = lambda dd: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in dd.items()}     # noqa: E731

    def loss_of(out):
        tot = 0.0
        if "mask" in case["heads"]:
            for c, m in zip(out["predictions_class"], out["predictions_mask"]):
                tot = tot + c.masked_fill(~torch.isfinite(c), 0.0).square().sum() * 0.01 + m.clamp_min(-100.0).square().sum() * 1e-4
        if "ground" in case["heads"]:
            lg = out["ground_logits"]
            tot = tot + lg.masked_fill(~torch.isfinite(lg), 0.0).square().sum()
        return tot
    torch.manual_seed(99)
    ref = model(clone(d))
    loss_of(ref).backward()

    class Hook(TorchRngTrain):
        def obj_dropout(self, name, x):
            return torch.nn.functional.dropout(x, 0.1, True)

        def head_dropout(self, head, h):
            return torch.nn.functional.dropout(h, 0.3 if head == "ground" else 0.1, True)
    ocfg = C.oracle_model_cfg(w, cfg)
    hook = Hook(0.1, 0.0, w.spatial_selfattn)
    ocfg.train = hook
    ocfg.decoder.train = hook
    sdd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("gauss_B") else v.clone())
           for k, v in sd.items()}
    torch.manual_seed(99)
    out = O.query3d_unified_forward(sdd, ocfg, clone(d))
    loss_of(out).backward()
    if "mask" in case["heads"]:
        a, b = out["predictions_mask"][-1], ref["predictions_mask"][-1]
        assert (a - b).abs().max() <= 1e-4 * b.abs().clamp_max(1e3).max()
    else:
        fin = torch.isfinite(ref["ground_logits"])
        assert torch.allclose(out["ground_logits"][fin], ref["ground_logits"][fin], atol=1e-5)
    gmax = max(float(p.grad.abs().max()) for _, p in model.named_parameters() if p.grad is not None)
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        gg = sdd[k].grad
        assert gg is not None, k
        assert float((gg - p.grad).abs().max()) <= 1e-4 * max(float(p.grad.abs().max()), 1e-3 * gmax), k
        n += 1
    assert n > 50
