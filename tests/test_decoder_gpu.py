"""GPU parity of the host-side mirror modules (running the sm_100a kernels through the C ABI)
against (1) the golden fixtures the REAL reference produced in fp32 and (2) the oracle restatement
run on the same GPU in fp32 and under torch.autocast(bf16) — the reference's own mixed-precision path.

Tolerance.  north_star asks for 1e-3 rel on bf16.  bf16 operands alone perturb a 768-long dot product
by ~1.6e-3 relative, so no bf16-operand implementation (the reference's autocast path included) is
within 1e-3 of the fp32 result end to end; what CAN be asserted, and is, is
  (a) err(ours, fp32) <= 1.25 * err(reference-under-autocast, fp32) + 1e-3   [||.||inf / ||.||inf]
      i.e. we are at least as close to the fp32 reference as the reference's bf16 path is, and
  (b) an absolute cap of 3e-2 on the same metric — waived only when the reference's own bf16 path is
      itself outside it (the deliberately ill-conditioned `sharp=4` fixture, whose near-one-hot softmax
      amplifies any operand rounding: there (a) alone applies),
with the measured values printed.  The 1e-3 bar itself is asserted where it is meaningful: per kernel,
on identical bf16-rounded operands, for fp32 outputs (tests/kernel_checks.py).
Bool masks: exact wherever the fp32 logit is not within rounding distance of the threshold.
"""
import pytest
import torch

import _cases as C
from oracle import restatement as O
from pq3d_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    fin = torch.isfinite(b) & (b.abs() < 1e5)
    return ((a[fin] - b[fin]).abs().max() / b[fin].abs().max().clamp_min(1e-12)).item()


def check_close(name, ours, gold, ref_bf16):
    e_ours, e_ref = rel(ours, gold), rel(ref_bf16, gold)
    print(f"{name}: err(ours, fp32 reference) = {e_ours:.3e}; err(autocast-bf16 oracle, fp32 reference) = {e_ref:.3e}")
    assert torch.isfinite(ours.float()).all() or not torch.isfinite(gold).all()
    assert e_ours <= 1.25 * e_ref + 1e-3, f"{name}: {e_ours:.3e} worse than the reference's own bf16 path {e_ref:.3e}"
    assert e_ours <= max(3e-2, 1.25 * e_ref), f"{name}: {e_ours:.3e} above the absolute cap"


def build_decoder(w, sd):
    from pq3d_b200.query_encoder import QueryMaskEncoder
    enc = QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
    enc.load_state_dict(sd, strict=True)
    return enc.to(DEV)


def oracle_autocast(fn):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return fn()


@pytest.mark.parametrize("name", C.golden_files("decoder"))
def test_decoder_vs_golden(name):
    g = C.load_golden(name)
    case = g["case"]
    w = C.build_workload(case)
    sd = synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"])
    enc = build_decoder(w, sd)
    inp, pw, _ = synth.make_decoder_inputs(w, device=DEV)
    with torch.no_grad():
        q, pc, pm = enc(synth.clone_input_dict(inp), pw)
    torch.cuda.synchronize()
    assert pc == [] and pm == []
    ref16 = oracle_autocast(lambda: O.query_mask_encoder(C.to_dev(sd, DEV), O.DecoderCfg(**w.decoder_kwargs()),
                                                         synth.clone_input_dict(inp), pw)[0])
    check_close(name, q, g["out"]["query"], ref16)


@pytest.mark.parametrize("name", C.golden_files("maskhead"))
def test_maskhead_vs_golden(name):
    from functools import partial
    from pq3d_b200.mask_head import MaskHeadSegLevel
    g = C.load_golden(name)
    case, gold = g["case"], g["out"]
    w = C.build_workload(case)
    sd = synth.decoder_state_dict(w, seed=case["wseed"], sharp=case["sharp"])
    n_match = len([m for m in w.memories if m in synth.SCENE_MEMORIES])
    sd_mh = synth.draw_state_dict(synth.mask_head_param_shapes(n_match), case["wseed"] + 100)
    enc = build_decoder(w, sd)
    mh = MaskHeadSegLevel(None, 768, 201, memories_for_match=list(w.memories), filter_out_classes=[0, 2]).eval()
    mh.load_state_dict(sd_mh, strict=True)
    mh = mh.to(DEV)
    inp, pw, d = synth.make_decoder_inputs(w, device=DEV)
    head = partial(mh, seg_fts_for_match=C.mask_head_inputs(w, inp), seg_masks=(~d["seg_pad_masks"]).to(DEV),
                   offline_attn_masks=None, skip_prediction=False)
    with torch.no_grad():
        q, pc, pm = enc(synth.clone_input_dict(inp), pw, head)
        c, m, a = head(query=q)
    torch.cuda.synchronize()
    assert len(pm) == int(gold["n_pred"])
    ref16 = oracle_autocast(lambda: C.oracle_maskhead(case, device=DEV))
    check_close(name + ".query", q, gold["query"], ref16["query"])
    check_close(name + ".pred_mask_first", pm[0], gold["pred_mask_first"], ref16["pred_mask_first"])
    check_close(name + ".final_mask", m, gold["final_mask"], ref16["final_mask"])
    check_close(name + ".final_class", c, gold["final_class"], ref16["final_class"])
    assert torch.equal(torch.isinf(c).cpu(), torch.isinf(gold["final_class"])), "filtered classes must be -inf"
    assert torch.equal((m == -1e6).cpu(), gold["final_mask"] == -1e6), "padded segments must be exactly -1e6"
    # first-call mask: inputs are identical (query = 0), so it differs from fp32 only by bf16 rounding of logits
    am0 = (pm[0].sigmoid().permute(0, 2, 1) < 0.5).cpu()
    gm0 = gold["pred_mask_first"].sigmoid().permute(0, 2, 1) < 0.5
    diff = am0 != gm0
    lg = gold["pred_mask_first"].permute(0, 2, 1)
    scale = lg[lg.abs() < 1e5].abs().max()
    assert (lg[diff].abs() <= 2e-2 * scale).all(), "attention-mask bits differ away from the decision threshold"
    print(f"{name}: first-call attn-mask bits differing from fp32 reference: {int(diff.sum())} / {diff.numel()} "
          f"(all within 2e-2*max|logit| of the threshold)")
    # the bool mask we return is exactly sigmoid(our logits) < 0.5
    assert torch.equal(a.cpu(), (m.sigmoid().permute(0, 2, 1) < 0.5).cpu())


@pytest.mark.parametrize("name", C.golden_files("model"))
def test_model_vs_golden(name):
    from pq3d_b200.query3d_unified import Query3DUnified
    g = C.load_golden(name)
    case, gold = g["case"], g["out"]
    w, cfg = C.model_cfg(case)
    sd = synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"])
    model = Query3DUnified(cfg).eval()
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV)
    d = C.to_dev(synth.make_model_data_dict(w, cfg), DEV)
    with torch.no_grad():
        out = model(d)
    torch.cuda.synchronize()
    ref16 = oracle_autocast(lambda: C.oracle_model(case, device=DEV))
    if "mask" in case["heads"]:
        assert len(out["predictions_mask"]) == int(gold["n_pred"])
        check_close(name + ".pred_mask_last", out["predictions_mask"][-1], gold["pred_mask_last"], ref16["pred_mask_last"])
        check_close(name + ".pred_class_last", out["predictions_class"][-1], gold["pred_class_last"], ref16["pred_class_last"])
    if "ground" in case["heads"]:
        check_close(name + ".ground_logits", out["ground_logits"], gold["ground_logits"], ref16["ground_logits"])
        assert torch.equal(torch.isinf(out["ground_logits"]).cpu(), torch.isinf(gold["ground_logits"]))


# ---- full-size, size-independent properties (BASELINE config 3 per-GPU shard) ------------------
@pytest.fixture(scope="module")
def c3():
    w = synth.workload("c3")
    sd = synth.decoder_state_dict(w, seed=0, sharp=2.0)
    enc = build_decoder(w, sd)
    inp, pw, d = synth.make_decoder_inputs(w, device=DEV)
    with torch.no_grad():
        base = enc(synth.clone_input_dict(inp), pw)[0]
    return w, sd, enc, inp, pw, base


def test_c3_matches_oracle_on_gpu(c3):
    w, sd, enc, inp, pw, base = c3
    sdd = C.to_dev(sd, DEV)
    cfg = O.DecoderCfg(**w.decoder_kwargs())
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref32 = O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw)[0]
    ref16 = oracle_autocast(lambda: O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw)[0])
    check_close("c3 (B=4,N=100,S=2048,[mv,pc,voxel,prompt],mixed)", base, ref32, ref16)


def test_c4_ragged_selfmask_matches_oracle_on_gpu():
    """BASELINE config 4 at full size: ragged scenes (S_b in [128, 4096]), 200 queries (two query tiles), in-loop mask
    head, per-query self masks with the all-masked-row fix-up, trailing padding tiles skipped — whole loop in one graph."""
    from functools import partial
    from pq3d_b200.mask_head import MaskHeadSegLevel
    w = synth.workload("c4")
    w.num_layers = 2                               # the mask feedback loop is chaotic; two layers keep bf16 noise bounded
    sd = synth.decoder_state_dict(w, seed=0, sharp=1.0)
    sd_mh = synth.draw_state_dict(synth.mask_head_param_shapes(3), 100)
    enc = build_decoder(w, sd)
    mh = MaskHeadSegLevel(None, 768, 201, memories_for_match=list(w.memories), filter_out_classes=[0, 2]).eval()
    mh.load_state_dict(sd_mh, strict=True)
    mh = mh.to(DEV)
    inp, pw, d = synth.make_decoder_inputs(w, device=DEV)
    seg_masks = (~d["seg_pad_masks"]).to(DEV)
    head = partial(mh, seg_fts_for_match=C.mask_head_inputs(w, inp), seg_masks=seg_masks, offline_attn_masks=None,
                   skip_prediction=False)
    outs = []
    with torch.no_grad():
        for _ in range(3):                         # eager, capture, replay
            q, pc, pm = enc(synth.clone_input_dict(inp), pw, head)
            outs.append((q, pm[0].clone(), pm[-1].clone()))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[1][2], outs[2][2]), "graph replay differs from eager"
    sdd, sdm = C.to_dev(sd, DEV), C.to_dev(sd_mh, DEV)
    cfg = O.DecoderCfg(**w.decoder_kwargs())
    ohead = partial(O.mask_head_seg_level, sd=sdm, prefix="", seg_fts_for_match=C.mask_head_inputs(w, inp),
                    seg_masks=seg_masks, filter_out_classes=[0, 2])
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        r32 = O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw, ohead)
    r16 = oracle_autocast(lambda: O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw, ohead))
    check_close("c4.pred_mask_first", outs[2][1], r32[2][0], r16[2][0])
    check_close("c4.query", outs[2][0], r32[0], r16[0])
    assert len(pm) == len(r32[2]) == 2


def test_c3_padding_invariance(c3):
    """Appending masked tokens must not change the output beyond tile-order rounding."""
    w, sd, enc, inp, pw, base = c3
    inp2 = synth.clone_input_dict(inp)
    g = torch.Generator().manual_seed(5)
    pad = 136
    newpos = None
    for m in ("mv", "pc", "voxel"):
        feat, mask, pos = inp2[m]
        B = feat.shape[0]
        if newpos is None:
            newpos = torch.cat([pos, torch.randn(B, pad, 768, generator=g).to(DEV)], 1)
        inp2[m] = [torch.cat([feat, torch.randn(B, pad, 768, generator=g).to(DEV)], 1),
                   torch.cat([mask, torch.ones(B, pad, dtype=torch.bool, device=DEV)], 1), newpos]
    with torch.no_grad():
        out = enc(inp2, pw)[0]
    e = rel(out, base)
    print(f"padding invariance: {e:.3e}")
    assert e <= 1e-3


def test_c3_permutation_equivariance(c3):
    w, sd, enc, inp, pw, base = c3
    inp2 = synth.clone_input_dict(inp)
    S = inp["mv"][0].shape[1]
    perm = torch.randperm(S, generator=torch.Generator().manual_seed(6)).to(DEV)
    for m in ("mv", "pc", "voxel"):
        feat, mask, pos = inp2[m]
        inp2[m] = [feat[:, perm].contiguous(), mask[:, perm].contiguous(), pos[:, perm].contiguous()]
    with torch.no_grad():
        out = enc(inp2, pw)[0]
    e = rel(out, base)
    print(f"permutation equivariance: {e:.3e}")
    assert e <= 1e-2          # bf16 rounding of P and of the PV partial sums is order dependent; measured 5.5e-3


def test_c3_idempotent_and_input_not_clobbered(c3):
    w, sd, enc, inp, pw, base = c3
    before = inp["mv"][0].clone()
    with torch.no_grad():
        again = enc(synth.clone_input_dict(inp), pw)[0]
    assert torch.equal(again, base), "same inputs must give bit-identical outputs"
    assert torch.equal(inp["mv"][0], before)


def test_all_keys_masked_gives_ln_of_bias():
    """Known answer: with every key masked only the zero-attn key is attended, so one sequential
    cross-attention layer returns LN(tgt + out_proj.bias) before self-attention/FFN; checked on the
    attention kernel directly: O must be exactly 0."""
    from pq3d_b200 import ops
    B, H, N, S = 2, 12, 100, 300
    D = H * 64
    Q = torch.randn(B * N, D, device=DEV).bfloat16()
    Sp = ops.pad8(S)
    K = torch.randn(B * Sp, D, device=DEV).bfloat16()
    Vt = torch.randn(D, B * Sp, device=DEV).bfloat16()
    bits = ops.pack_mask(torch.ones(B, S, dtype=torch.bool, device=DEV))
    O_ = torch.full((1, B * N, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    ops.attention(Q, 0, [ops.AttnMemory(K, 0, Vt, 0, S, Sp, bits, bits.stride(0), 0, 0)], O_, B * N * D, B, H, N, True)
    torch.cuda.synchronize()
    assert (O_ == 0).all()


def test_rejects_non_bool_masks_and_training_mode():
    w = synth.Workload("t", 1, 16, 64, ["pc"], "parallel", num_layers=1)
    enc = build_decoder(w, synth.decoder_state_dict(w, seed=1))
    inp, pw, _ = synth.make_decoder_inputs(w, device=DEV)
    bad = synth.clone_input_dict(inp)
    bad["pc"][1] = bad["pc"][1].float()
    with torch.no_grad(), pytest.raises(TypeError):
        enc(bad, pw)
    with pytest.raises(NotImplementedError):
        enc(synth.clone_input_dict(inp), pw, lambda q: (None, None, None))   # grad enabled + mask head: no backward
    enc.train()                                           # module.train(): the reference's dropout 0.1 is active
    with torch.no_grad():
        a = enc(synth.clone_input_dict(inp), pw)[0]
        b = enc(synth.clone_input_dict(inp), pw)[0]
    assert torch.isfinite(a).all() and not torch.equal(a, b)     # two draws of the dropout masks
    with pytest.raises(NotImplementedError):
        enc(synth.clone_input_dict(inp), pw, lambda q: (None, None, None))   # train mode + mask head: no backward


def test_whole_forward_graph_on_stable_buffers_and_concurrent_streams():
    """Serving-loop paths: (1) the same device tensors handed in again -> the prologue is captured with the body (one
    graph launch per forward) and still reads the buffers' CURRENT contents; (2) two batches in flight on two CUDA
    streams use separate workspaces.  Both must reproduce the plain eager-prologue result bit for bit."""
    w = synth.Workload("tfull", 2, 100, 384, ["mv", "pc", "voxel", "prompt"], "mixed", T=12, num_layers=2)
    sd = synth.decoder_state_dict(w, seed=3)
    enc = build_decoder(w, sd)
    inp_a, pw, _ = synth.make_decoder_inputs(w, device=DEV)
    w2 = synth.Workload("tfull", 2, 100, 384, ["mv", "pc", "voxel", "prompt"], "mixed", T=12, num_layers=2, seed=999)
    inp_b, pw_b, _ = synth.make_decoder_inputs(w2, device=DEV)
    enc.use_cuda_graph = False
    with torch.no_grad():
        ref_a = enc(synth.clone_input_dict(inp_a), pw)[0].clone()
        ref_b = enc(synth.clone_input_dict(inp_b), pw_b)[0].clone()
    assert not torch.equal(ref_a, ref_b)
    enc.use_cuda_graph = True
    enc._ws.clear()
    with torch.no_grad():
        outs = [enc(synth.clone_input_dict(inp_a), pw)[0].clone() for _ in range(6)]     # eager, body graph, full graph
    ws = next(iter(enc._ws.values()))
    assert any(e.get("graph") is not None for e in ws.get("full", {}).values()), "whole-forward graph was not captured"
    for o in outs:
        assert torch.equal(o, ref_a)
    # new contents written into the SAME buffers are picked up by the replay
    flat_a = [t for v in inp_a.values() for t in enc._flat_tensors(list(v))]
    flat_b = [t for v in inp_b.values() for t in enc._flat_tensors(list(v))]
    with torch.no_grad():
        for ta, tb in zip(flat_a, flat_b):
            ta.copy_(tb)
        pw.copy_(pw_b)
        assert torch.equal(enc(synth.clone_input_dict(inp_a), pw)[0], ref_b)
    # two streams, two batches in flight
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    inp_c, pw_c, _ = synth.make_decoder_inputs(w, device=DEV)
    res = {}
    with torch.no_grad():
        for _ in range(5):
            for st, (name, i_, p_) in ((s1, ("a", inp_c, pw_c)), (s2, ("b", inp_a, pw))):
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    res[name] = enc(synth.clone_input_dict(i_), p_)[0]
    torch.cuda.synchronize()
    assert torch.equal(res["a"], ref_a) and torch.equal(res["b"], ref_b)
    assert len(enc._ws) >= 3


@pytest.mark.parametrize("dim_loc", [3, 6])
def test_location_prompts(dim_loc):
    """prompt_encoder's location branch (model/query3d_unified.py:95-102) through the coordinate-encoder kernels."""
    import contextlib
    from _train_hooks import run_prompt_loc_case
    run_prompt_loc_case(dim_loc, DEV, contextlib.nullcontext())


def test_model_producers_emit_decoder_operands_bit_identical(monkeypatch):
    """in inference the ObjectEncoder LayerNorm epilogue writes the decoder's bf16 K / V operands
    (xv = bf16(feat), xk = bf16(feat + fts_pos)) and the decoder skips its ingest pass.  Same rounding points as the
    ingest kernel, so the model output must be BIT-IDENTICAL with the path switched off, and no ingest may be launched."""
    from pq3d_b200 import ops
    from pq3d_b200.query3d_unified import Query3DUnified
    w = synth.Workload("pre", 2, 40, 256, ["mv", "pc", "voxel", "prompt"], "mixed", T=8, num_layers=2)
    cfg = synth.model_cfg_dict(w, dim_loc=3, heads=("ground",))
    model = Query3DUnified(cfg).eval()
    model.load_state_dict(synth.draw_state_dict(synth.model_param_shapes(cfg), 5, 1.5), strict=True)
    model = model.to(DEV)
    d = C.to_dev(synth.make_model_data_dict(w, cfg), DEV)
    calls = {"n": 0}
    real_many, real_one = ops.ingest_memories, ops.ingest_memory

    def count_many(*a, **k):
        calls["n"] += 1
        return real_many(*a, **k)

    def count_one(feat, *a, **k):
        calls["n"] += 1 if feat.shape[1] == 256 else 0          # the scene memories (the prompt keeps its own ingest)
        return real_one(feat, *a, **k)
    monkeypatch.setattr(ops, "ingest_memories", count_many)
    monkeypatch.setattr(ops, "ingest_memory", count_one)
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PQ3D_PREINGEST", flag)
        calls["n"] = 0
        with torch.no_grad():
            for _ in range(3):                                   # eager, captured, replayed decoder body
                outs[flag] = model(dict(d))["ground_logits"].clone()
        torch.cuda.synchronize()
        assert (calls["n"] == 0) == (flag == "1"), f"PQ3D_PREINGEST={flag}: {calls['n']} scene-memory ingest launches"
    assert torch.equal(outs["1"], outs["0"])
    assert torch.isfinite(outs["1"][d["query_pad_masks"]]).all()


def test_model_whole_forward_graph_replays_current_buffer_contents():
    """Query3DUnified.forward captures the WHOLE model forward into one CUDA graph when a serving loop hands in the same
    device tensors again: results bit-identical to the eager path, the replay reads the buffers' current contents, fresh
    tensors fall back to the eager path, and outputs are fresh tensors (never views of the graph's static memory)."""
    from pq3d_b200.query3d_unified import Query3DUnified
    w = synth.Workload("mgraph", 2, 40, 256, ["mv", "pc", "voxel", "prompt"], "mixed", T=8, num_layers=2)
    cfg = synth.model_cfg_dict(w, dim_loc=3, heads=("ground",))
    model = Query3DUnified(cfg).eval()
    model.load_state_dict(synth.draw_state_dict(synth.model_param_shapes(cfg), 7, 1.5), strict=True)
    model = model.to(DEV)
    d = C.to_dev(synth.make_model_data_dict(w, cfg), DEV)
    valid = d["query_pad_masks"]

    def eager(dd):
        model.use_cuda_graph = False
        try:
            with torch.no_grad():
                return model(dict(dd))["ground_logits"].clone()
        finally:
            model.use_cuda_graph = True
    ref1 = eager(d)
    with torch.no_grad():
        outs = [model(dict(d))["ground_logits"] for _ in range(4)]       # eager, capture + replay, replay, replay
    torch.cuda.synchronize()
    ents = [e for e in model._graphs.values() if e.get("graph") is not None]
    assert len(ents) == 1 and ents[0]["launches"] > 50
    for o in outs:
        assert torch.equal(o, ref1)
    assert len({o.data_ptr() for o in outs}) == len(outs)
    # refill the staging buffers in place with another batch: the replay must see it
    d2 = C.to_dev(synth.make_model_data_dict(w, cfg, rank=1), DEV)
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v.copy_(d2[k])
    with torch.no_grad():
        out2 = model(dict(d))["ground_logits"]
    ref2 = eager(d)
    assert torch.equal(out2, ref2)
    assert not torch.equal(ref2[valid], ref1[valid])
    assert torch.equal(outs[-1], ref1)                                   # earlier results were not overwritten
    # different tensors of the same shapes: eager path (no stale replay), same numbers
    with torch.no_grad():
        out3 = model(dict(d2))["ground_logits"]
    assert torch.equal(out3, ref2)
    # autograd / training never take the graph
    with torch.enable_grad():
        assert model._forward_graphed(dict(d)) is None
    # parameters updated in place (optimizer step / load_state_dict): the graphs are dropped, the next calls run the new
    # weights eagerly and capture again
    with torch.no_grad():
        for _ in range(2):
            model(dict(d))
        assert any(e.get("graph") is not None for e in model._graphs.values())
        model.ground_head.og3d_head[0].weight.mul_(1.5)
        model.mv_encoder.input_feat_proj[0].weight.mul_(0.5)
        model.unified_encoder.unified_encoder[0].ffn.linear1.weight.mul_(0.7)
        got = [model(dict(d))["ground_logits"] for _ in range(3)]         # eager, capture + replay, replay
    ref3 = eager(d)
    assert not torch.equal(ref3[valid], ref2[valid])
    for o in got:
        assert torch.equal(o, ref3)
