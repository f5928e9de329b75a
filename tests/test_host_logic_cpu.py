"""CPU: the product's HOST logic (operand layouts, grouped-launch arguments, packed-weight bookkeeping, gradient routing
of the training engine) run end to end against tests/_cpu_ops.py — a torch emulation of the kernels written from the
ABI documentation — and compared with the oracle.  No kernel runs here; kernel numerics are the `-m gpu` suite's job.
Tolerances are bf16-level (the emulation rounds where the kernels round)."""
import pytest
import torch

import _cpu_ops
from oracle import restatement as O
from pq3d_b200 import synth
from pq3d_b200.query_encoder import QueryMaskEncoder


def rel(a, b, floor=1e-20):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(floor)).item()


def _case(structure="mixed", spatial=True, N=24, S=150, L=2):
    mems = ["mv", "pc", "voxel"] + (["prompt"] if structure != "parallel" else [])
    w = synth.Workload("h", 2, N, S, mems, structure, T=9, num_layers=L, spatial_selfattn=spatial, ragged=(S // 2, S))
    sd = synth.decoder_state_dict(w, seed=11, sharp=1.0)
    inp, pw, _ = synth.make_decoder_inputs(w)
    g = torch.Generator().manual_seed(5)
    q, qm, qp = inp["query"]
    inp["query"] = (torch.randn(q.shape, generator=g) * 0.5, qm, qp)
    return w, sd, inp, pw, torch.randn(q.shape, generator=g)


def xo_mask_expected(rep, H):
    """(B*H, N, S) built by repeat_interleave(H, 0): rows b*H .. b*H+H-1 are identical copies."""
    return rep.view(-1, H, *rep.shape[1:])[:, :1].expand(-1, H, -1, -1).reshape(rep.shape)


def _build(w, sd, **kw):
    enc = QueryMaskEncoder(None, **dict(w.decoder_kwargs(), **kw))
    enc.load_state_dict(sd, strict=True)
    enc.use_cuda_graph = False
    enc.train_streams = False
    return enc


@pytest.mark.parametrize("structure,spatial", [("mixed", True), ("sequential", False), ("parallel", True)])
def test_inference_host_logic(structure, spatial):
    w, sd, inp, pw, _ = _case(structure, spatial)
    enc = _build(w, sd).eval()
    with _cpu_ops.cpu_backend(), torch.no_grad():
        out = enc(synth.clone_input_dict(inp), pw)[0]
    ref = O.query_mask_encoder(sd, O.DecoderCfg(**w.decoder_kwargs()), synth.clone_input_dict(inp), pw)[0]
    assert rel(out, ref) <= 3e-2


@pytest.mark.parametrize("structure,spatial", [("mixed", True), ("sequential", False), ("gate", True)])
def test_training_host_logic_gradients(structure, spatial):
    w, sd, inp, pw, up = _case(structure, spatial)
    enc = _build(w, sd).train()
    enc.train_dropout = 0.0
    leaves = {}

    def leaf(t, name):
        leaves[name] = t.clone().requires_grad_(True)
        return leaves[name]
    x = {}
    q, qm, qp = inp["query"]
    x["query"] = (leaf(q, "query"), qm, leaf(qp, "query_pos"))
    pos_leaf = {}
    for m, (feat, mask, pos) in ((k, v) for k, v in inp.items() if k != "query"):
        p = None
        if pos is not None:
            if id(pos) not in pos_leaf:
                pos_leaf[id(pos)] = leaf(pos, f"pos[{m}]")
            p = pos_leaf[id(pos)]
        x[m] = [leaf(feat, f"{m}.feat"), mask, p]
    with _cpu_ops.cpu_backend():
        out = enc(x, pw)[0]
        (out * up).sum().backward()
    ours = {k: p.grad for k, p in enc.named_parameters()}
    ours.update({k: v.grad for k, v in leaves.items()})
    # oracle autograd
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
    xo = {"query": (lo["query"], qm, lo["query_pos"])}
    for m, (feat, mask, pos) in ((k, v) for k, v in inp.items() if k != "query"):
        xo[m] = [lo[f"{m}.feat"], mask, None if pos is None else lo[[k for k in lo if k.startswith("pos[")][0]]]
    def oracle(autocast):
        sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lo = {k: v.detach().clone().requires_grad_(True) for k, v in leaves.items()}
        xo = {"query": (lo["query"], qm, lo["query_pos"])}
        for m, (feat, mask, pos) in ((k, v) for k, v in inp.items() if k != "query"):
            xo[m] = [lo[f"{m}.feat"], mask, None if pos is None else lo[[k for k in lo if k.startswith("pos[")][0]]]
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            r = O.query_mask_encoder(sdd, O.DecoderCfg(**w.decoder_kwargs()), xo, pw)[0]
        (r.float() * up).sum().backward()
        g = {k: v.grad for k, v in sdd.items()}
        g.update({k: v.grad for k, v in lo.items()})
        return r.detach().float(), g
    ref, g32 = oracle(False)
    ref16, g16 = oracle(True)          # the reference's own bf16 path is the yardstick, as in tests/test_train_gpu.py
    assert rel(out, ref) <= 1.25 * rel(ref16, ref) + 2e-3
    floor = 1e-3 * max(float(v.abs().max()) for k, v in g32.items() if v is not None and k in sd)
    for k, r in g32.items():
        if r is None or k.endswith("w_ks.bias"):
            continue
        assert ours.get(k) is not None, f"{k}: no gradient"
        assert ours[k].shape == r.shape
        # Frobenius-norm error: with only 48 query rows a single ReLU gate flipped by bf16 rounding moves the max-norm
        # of an FFN weight gradient by tens of percent (in the autocast oracle too); the 2-norm sees the routing bugs
        # this test is after (a wrong slice, stride or scale is an O(1) error) without that brittleness
        l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(floor)).item()     # noqa: E731
        e, e16 = l2(ours[k], r), l2(g16[k], r)
        assert e <= 1.5 * e16 + 2e-2, f"{k}: {e:.3e} (autocast oracle {e16:.3e})"


def test_training_host_logic_dropout_and_memory_dropout():
    """Train mode proper on the emulated kernels: the oracle replays the counter-RNG masks (tests/_train_hooks.py), so
    forward and gradients must agree — checks site numbering, element indexing and the memory-dropout weights end to end
    on the host side."""
    from _train_hooks import KernelRngTrain
    w, sd, inp, pw, up = _case("mixed", True)
    enc = _build(w, sd, memory_dropout=0.6).train()
    enc.train_dropout = 0.1
    torch.manual_seed(3)
    with _cpu_ops.cpu_backend():
        x = synth.clone_input_dict(inp)
        q, qm, qp = x["query"]
        x["query"] = (q.clone().requires_grad_(True), qm, qp)
        out = enc(x, pw)[0]
        (out * up).sum().backward()
    seed = int(enc._drop_seed.item()) & 0xFFFFFFFF
    kw = dict(w.decoder_kwargs(), memory_dropout=0.6)
    cfg = O.DecoderCfg(**kw)
    cfg.train = KernelRngTrain(enc, seed, 0.1, w.B, w.N, w.num_heads)
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw)[0]
    (ref * up).sum().backward()
    assert rel(out, ref) <= 3e-2
    for k, p in enc.named_parameters():
        if k.endswith("w_ks.bias"):
            continue
        r = sdd[k].grad
        e = ((p.grad - r).norm() / r.norm().clamp_min(1e-12)).item()
        assert e <= 0.12, f"{k}: {e:.3e}"


def _mask_head_setup(B=2, N=20, S=90, n=2, seed=8):
    from pq3d_b200.mask_head import MaskHeadSegLevel
    g = torch.Generator().manual_seed(seed)
    mems = ["voxel", "mv", "pc"][:n]
    mh = MaskHeadSegLevel(None, 768, 201, memories_for_match=mems, filter_out_classes=[0, 2], dropout=0.0)
    sd = synth.draw_state_dict(synth.mask_head_param_shapes(n), seed)
    mh.load_state_dict(sd, strict=True)
    feats = []
    for j in range(n):
        mask = torch.rand(B, S, generator=g) < 0.2
        feats.append([torch.randn(B, S, 768, generator=g), mask, None])
    seg_masks = torch.zeros(B, S, dtype=torch.bool)
    seg_masks[1, S - 7:] = True
    for f in feats:
        f[1][1, S - 7:] = True
    q = torch.randn(B, N, 768, generator=g) * 0.5
    return mh, sd, feats, seg_masks, q, g


def test_mask_head_training_host_logic():
    """MaskHeadSegLevel.forward under autograd (the call Query3DUnified makes after the decoder) on the emulated
    kernels: predictions, query gradient, feature gradients and every parameter gradient against oracle autograd."""
    mh, sd, feats, seg_masks, q, g = _mask_head_setup()
    mh.train()
    B, N, _ = q.shape
    S = seg_masks.shape[1]
    up_c, up_m = torch.randn(B, N, 201, generator=g), torch.randn(B, S, N, generator=g)
    ql = q.clone().requires_grad_(True)
    fl = [[f[0].clone().requires_grad_(True), f[1], None] for f in feats]
    with _cpu_ops.cpu_backend():
        cls, ml, attn = mh(ql, fl, seg_masks)
        fin = torch.isfinite(cls)
        ((cls.masked_fill(~fin, 0.0) * up_c).sum() + (ml * up_m).sum()).backward()
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    qo = q.clone().requires_grad_(True)
    fo = [[f[0].clone().requires_grad_(True), f[1], None] for f in feats]
    cls_o, ml_o, attn_o = O.mask_head_seg_level(qo, sdd, "", fo, seg_masks, filter_out_classes=[0, 2])
    ((cls_o.masked_fill(~torch.isfinite(cls_o), 0.0) * up_c).sum() + (ml_o * up_m).sum()).backward()
    assert torch.equal(torch.isfinite(cls), torch.isfinite(cls_o))
    assert rel(cls.masked_fill(~fin, 0.0), cls_o.masked_fill(~fin, 0.0)) <= 2e-2
    assert rel(ml, ml_o) <= 2e-2
    l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()     # noqa: E731
    assert l2(ql.grad, qo.grad) <= 5e-2
    for a, b in zip(fl, fo):
        assert l2(a[0].grad, b[0].grad) <= 5e-2
    for k, p in mh.named_parameters():
        assert p.grad is not None, k
        assert l2(p.grad, sdd[k].grad) <= 5e-2, (k, l2(p.grad, sdd[k].grad))


def test_stage1_training_host_logic_mask_head_selfmask_blocks():
    """Stage-1 (instance segmentation) training shape on the emulated kernels: parallel cross-attentions, the in-loop
    mask head whose detached attention mask replaces the memory masks (use_self_mask), two blocks re-applying the same
    layers, losses on every per-layer prediction — gradients of decoder and mask-head parameters vs oracle autograd."""
    from functools import partial
    from pq3d_b200.mask_head import MaskHeadSegLevel
    B, N, S, L, K = 2, 16, 80, 2, 2
    w = synth.Workload("s1", B, N, S, ["mv", "pc", "voxel"], "parallel", num_layers=L, num_blocks=K, use_self_mask=True,
                       spatial_selfattn=True)
    sd = synth.decoder_state_dict(w, seed=13, sharp=1.0)
    inp, pw, dd = synth.make_decoder_inputs(w)
    g = torch.Generator().manual_seed(21)
    q, qm, qp = inp["query"]
    inp["query"] = (torch.randn(q.shape, generator=g) * 0.5, qm, qp)
    msd = synth.draw_state_dict(synth.mask_head_param_shapes(3), 17)
    mh = MaskHeadSegLevel(None, 768, 201, memories_for_match=list(w.memories), filter_out_classes=[0, 2], dropout=0.0)
    mh.load_state_dict(msd, strict=True)
    mh.train()
    seg_pad = ~dd["seg_pad_masks"]
    enc = _build(w, sd).train()
    enc.train_dropout = 0.0
    n_pred = K * L
    ups = [(torch.randn(B, N, 201, generator=g), torch.randn(B, S, N, generator=g)) for _ in range(n_pred)]
    up_q = torch.randn(B, N, 768, generator=g)

    def loss_of(out, pcs, pms):
        tot = (out * up_q).sum()
        for (uc, um), c, m in zip(ups, pcs, pms):
            tot = tot + (c.masked_fill(~torch.isfinite(c), 0.0) * uc).sum() * 0.1 + (m.clamp_min(-100.0) * um).sum() * 0.1
        return tot
    x = synth.clone_input_dict(inp)
    feats = [list(x[m]) for m in w.memories]
    head = partial(mh, seg_fts_for_match=feats, seg_masks=seg_pad, offline_attn_masks=None, skip_prediction=False)
    with _cpu_ops.cpu_backend():
        out, pcs, pms = enc(x, pw, head)
        assert len(pcs) == n_pred and len(pms) == n_pred
        loss_of(out, pcs, pms).backward()
    assert x["mv"][1].shape == (B * enc.num_heads, N, S) and x["mv"][1].dtype == torch.bool   # reference side effect (:84-88)
    assert torch.equal(x["mv"][1], xo_mask_expected(x["mv"][1], enc.num_heads))
    # oracle
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    msdd = {k: v.clone().requires_grad_(True) for k, v in msd.items()}
    xo = synth.clone_input_dict(inp)
    feats_o = [list(xo[m]) for m in w.memories]
    head_o = lambda qq: O.mask_head_seg_level(qq, msdd, "", feats_o, seg_pad, filter_out_classes=[0, 2])   # noqa: E731
    ro, pco, pmo = O.query_mask_encoder(sdd, O.DecoderCfg(**w.decoder_kwargs()), xo, pw, head_o)
    loss_of(ro, pco, pmo).backward()
    flips = sum(int(((a < 0) != (b < 0)).sum()) for a, b in zip(pms, pmo))
    print(f"mask-logit sign flips vs oracle: {flips} of {sum(a.numel() for a in pms)}")
    assert rel(out, ro) <= 5e-2
    l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()     # noqa: E731
    worst = []
    for k, p in enc.named_parameters():
        if k.endswith("w_ks.bias"):
            continue
        assert p.grad is not None, k
        worst.append((l2(p.grad, sdd[k].grad), k))
    for k, p in mh.named_parameters():
        assert p.grad is not None, k
        worst.append((l2(p.grad, msdd[k].grad), "mask_head." + k))
    worst.sort(reverse=True)
    print(worst[:6])
    assert worst[0][0] <= 0.15, worst[:4]


def test_training_host_logic_multiscale_voxel():
    """Multi-scale voxel memory (a list with one feature table per layer, query_encoder.py:90-91) in training: every
    scale's feature gradient and the per-layer K / V weight gradients against oracle autograd."""
    w = synth.Workload("ms", 2, 20, 100, ["voxel", "pc"], "parallel", num_layers=2, voxel_multiscale=True,
                       spatial_selfattn=False)
    sd = synth.decoder_state_dict(w, seed=19, sharp=1.0)
    inp, pw, _ = synth.make_decoder_inputs(w)
    assert isinstance(inp["voxel"][0], list)
    g = torch.Generator().manual_seed(2)
    up = torch.randn(inp["query"][0].shape, generator=g)
    enc = _build(w, sd).train()
    enc.train_dropout = 0.0
    x = synth.clone_input_dict(inp)
    vox = [t.clone().requires_grad_(True) for t in x["voxel"][0]]
    x["voxel"][0] = list(vox)
    with _cpu_ops.cpu_backend():
        out = enc(x, pw)[0]
        (out * up).sum().backward()
    assert torch.is_tensor(x["voxel"][0])
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = synth.clone_input_dict(inp)
    vo = [t.clone().requires_grad_(True) for t in xo["voxel"][0]]
    xo["voxel"][0] = list(vo)
    ref = O.query_mask_encoder(sdd, O.DecoderCfg(**w.decoder_kwargs()), xo, pw)[0]
    (ref * up).sum().backward()
    assert rel(out, ref) <= 3e-2
    l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()     # noqa: E731
    assert sum(b.grad is not None for b in vo) == w.num_layers          # layer i attends scale i; the rest are unused
    for a, b in zip(vox, vo):
        if b.grad is None:
            assert a.grad is None
        else:
            assert a.grad is not None and l2(a.grad, b.grad) <= 8e-2
    for k, p in enc.named_parameters():
        assert p.grad is not None, k
        assert l2(p.grad, sdd[k].grad) <= 0.12, (k, l2(p.grad, sdd[k].grad))


@pytest.mark.parametrize("stage", ["stage1_mask", "stage2_ground"])
def test_model_level_training_host_logic(stage):
    """Query3DUnified in .train() on the emulated kernels, dropouts off: gradients of EVERY model parameter (object
    encoders, coordinate encoder(s), decoder, mask head / ground head) against autograd through the oracle's
    query3d_unified_forward — the trainer-facing boundary (`loss.backward()` after `model(data_dict)`)."""
    from _train_hooks import run_model_training_case
    run_model_training_case(stage, "cpu", _cpu_ops.cpu_backend())


@pytest.mark.parametrize("dim_loc", [3, 6])
def test_location_prompts_host_logic(dim_loc):
    from _train_hooks import run_prompt_loc_case
    run_prompt_loc_case(dim_loc, "cpu", _cpu_ops.cpu_backend())


def test_training_host_logic_share_layer():
    """share_layer=True (one QueryEncoderLayer object used by every layer, modules/utils.py:28-32): the gradient of each
    shared parameter is the sum over the layers that use it."""
    w, sd, inp, pw, up = _case("mixed", True)
    L = w.num_layers
    for k in list(sd):                                   # every layer carries layer 0's weights
        if k.startswith("unified_encoder.0."):
            for l in range(1, L):
                sd[k.replace("unified_encoder.0.", f"unified_encoder.{l}.")] = sd[k].clone()
    enc = _build(w, sd, share_layer=True).train()
    enc.train_dropout = 0.0
    with _cpu_ops.cpu_backend():
        out = enc(synth.clone_input_dict(inp), pw)[0]
        (out * up).sum().backward()
    sdd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.query_mask_encoder(sdd, O.DecoderCfg(**w.decoder_kwargs()), synth.clone_input_dict(inp), pw)[0]
    (ref * up).sum().backward()
    assert rel(out, ref) <= 3e-2
    named = dict(enc.named_parameters())                 # de-duplicated: names of layer 0
    assert all(k.startswith("unified_encoder.0.") for k in named)
    for k, p in named.items():
        if k.endswith("w_ks.bias"):
            continue
        r = sum(sdd[k.replace("unified_encoder.0.", f"unified_encoder.{l}.")].grad for l in range(L))
        e = ((p.grad - r).norm() / r.norm().clamp_min(1e-12)).item()
        assert e <= 0.12, f"{k}: {e:.3e}"


def test_inference_host_logic_in_loop_mask_head_selfmask_blocks():
    """Inference with our MaskHeadSegLevel in the loop (the path that goes into ONE CUDA graph on the GPU: hoisted
    k-projection, per-layer head + mask packing + layer on static buffers), use_self_mask, two blocks, on the emulated
    kernels: per-layer predictions and the decoded queries vs the oracle; the attention-mask side effect on input_dict."""
    from functools import partial
    from pq3d_b200.mask_head import MaskHeadSegLevel
    B, N, S, L, K = 2, 16, 80, 2, 2
    w = synth.Workload("i1", B, N, S, ["mv", "pc", "voxel"], "parallel", num_layers=L, num_blocks=K, use_self_mask=True)
    sd = synth.decoder_state_dict(w, seed=23, sharp=1.0)
    inp, pw, dd = synth.make_decoder_inputs(w)
    msd = synth.draw_state_dict(synth.mask_head_param_shapes(3), 27)
    mh = MaskHeadSegLevel(None, 768, 201, memories_for_match=list(w.memories), filter_out_classes=[0, 2]).eval()
    mh.load_state_dict(msd, strict=True)
    seg_pad = ~dd["seg_pad_masks"]
    enc = _build(w, sd).eval()
    x = synth.clone_input_dict(inp)
    head = partial(mh, seg_fts_for_match=[list(x[m]) for m in w.memories], seg_masks=seg_pad, offline_attn_masks=None,
                   skip_prediction=False)
    with _cpu_ops.cpu_backend(), torch.no_grad():
        out, pcs, pms = enc(x, pw, head)
    xo = synth.clone_input_dict(inp)
    fo = [list(xo[m]) for m in w.memories]
    with torch.no_grad():
        ro, pco, pmo = O.query_mask_encoder(sd, O.DecoderCfg(**w.decoder_kwargs()), xo, pw,
                                            lambda q: O.mask_head_seg_level(q, msd, "", fo, seg_pad, filter_out_classes=[0, 2]))
    assert len(pcs) == len(pco) == K * L
    assert rel(out, ro) <= 5e-2
    for a, b in zip(pms, pmo):
        keep = b > -1e5
        assert rel(a[keep], b[keep]) <= 5e-2
    for a, b in zip(pcs, pco):
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin) and rel(a[fin], b[fin]) <= 5e-2
    assert x["mv"][1].dtype == torch.bool and x["mv"][1].shape == (B * enc.num_heads, N, S)
    assert x["mv"][1].shape == xo["mv"][1].shape                      # same layout as the oracle / reference leaves


def test_inference_host_logic_query_encoder_and_dropped_memories():
    """QueryEncoder (the mask-less class) and drop_memories_test in eval on the emulated kernels vs the oracle."""
    from pq3d_b200.query_encoder import QueryEncoder
    w, sd, inp, pw, _ = _case("mixed", True)
    enc = _build(w, sd, drop_memories_test=["pc"]).eval()
    with _cpu_ops.cpu_backend(), torch.no_grad():
        out = enc(synth.clone_input_dict(inp), pw)[0]
    cfg = O.DecoderCfg(**dict(w.decoder_kwargs(), drop_memories_test=["pc"]))
    ref = O.query_mask_encoder(sd, cfg, synth.clone_input_dict(inp), pw)[0]
    assert rel(out, ref) <= 3e-2
    kw = {k: v for k, v in w.decoder_kwargs().items() if k not in ("use_self_mask", "num_blocks")}
    qe = QueryEncoder(None, **kw)
    qe.load_state_dict(sd, strict=True)
    qe.use_cuda_graph = False
    qe.eval()
    with _cpu_ops.cpu_backend(), torch.no_grad():
        out2 = qe(synth.clone_input_dict(inp), pw)
    ref2 = O.query_mask_encoder(sd, O.DecoderCfg(**w.decoder_kwargs()), synth.clone_input_dict(inp), pw)[0]
    assert rel(out2, ref2) <= 3e-2


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="/root/reference not present")
def test_query_encoder_train_mode_memory_dropout_matches_live_reference():
    """QueryEncoder.train() (query_encoder.py:26-36): each scene's feat / pos of every scene memory is zeroed with
    probability memory_dropout (torch.rand per memory, in place), the layers themselves drop nothing.  Same torch seed
    -> same draws as the LIVE reference class; outputs compared with the reference's sublayer dropouts switched off."""
    from oracle import ref_loader
    from pq3d_b200.query_encoder import QueryEncoder
    ns = ref_loader.load()
    w, sd, inp, pw, _ = _case("mixed", True)
    kw = {k: v for k, v in w.decoder_kwargs().items() if k not in ("use_self_mask", "num_blocks")}
    kw["memory_dropout"] = 0.5
    ref = ns.query_encoder.QueryEncoder(None, **kw).train()
    ref.load_state_dict(sd, strict=True)
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    ours = QueryEncoder(None, **kw).train()
    ours.load_state_dict(sd, strict=True)
    ours.use_cuda_graph, ours.train_streams, ours.train_dropout = False, False, 0.0
    assert ours.memory_dropout == 0.5 and all(l.memory_dropout == 0 for l in ours.unified_encoder)
    hit = False
    for seed in (3, 4, 5):
        xr, xo = synth.clone_input_dict(inp), synth.clone_input_dict(inp)
        torch.manual_seed(seed)
        with torch.no_grad():
            r = ref(xr, pw)
        torch.manual_seed(seed)
        with _cpu_ops.cpu_backend(), torch.no_grad():
            o = ours(xo, pw)
        for m in ("mv", "pc", "voxel"):                     # the caller's tensors are zeroed in place, identically
            assert torch.equal(xr[m][0], xo[m][0]) and torch.equal(xr[m][2], xo[m][2])
            hit = hit or bool(ours.last_scene_drop[m].any())
        assert rel(o, r) <= 3e-2
    assert hit, "no scene was dropped in three draws at p = 0.5"


def test_dropout_sites_are_distinct_streams():
    """Site numbering of the counter RNG: every dropout call site of every layer (and every mask-head call) gets its own
    stream, for up to 8 memories and 64 layers; masks of different sites are decorrelated."""
    from pq3d_b200 import rng
    sites = set()
    for layer in range(64):
        kinds = [rng.SITE_CA_SUBLAYER + g for g in range(4)] + [rng.SITE_SA_SUBLAYER, rng.SITE_FFN_SUBLAYER,
                                                                rng.SITE_FFN_HIDDEN, rng.SITE_SA_PROBS]
        kinds += [rng.SITE_CA_PROBS + j for j in range(8)]
        assert max(kinds) < rng.SITES_PER_LAYER and len(set(kinds)) == len(kinds)
        for k in kinds:
            sites.add(rng.site(layer, k))
    assert len(sites) == 64 * 16 and max(sites) < rng.SITE_MASK_HEAD
    idx = torch.arange(1 << 14)
    a = rng.keep_mask(123, rng.site(0, rng.SITE_FFN_HIDDEN), idx, 0.5)
    b = rng.keep_mask(123, rng.site(1, rng.SITE_FFN_HIDDEN), idx, 0.5)
    c = rng.keep_mask(124, rng.site(0, rng.SITE_FFN_HIDDEN), idx, 0.5)
    for other in (b, c):
        agree = (a == other).float().mean().item()
        assert 0.45 < agree < 0.55
