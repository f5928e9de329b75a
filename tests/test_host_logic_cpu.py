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


@pytest.mark.parametrize("structure,spatial", [("mixed", True), ("sequential", False)])
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
