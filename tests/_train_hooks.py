"""Shared by the GPU and the CPU (emulated-kernel) training tests: the oracle hook that replays the kernels' dropout
masks."""
import torch


class KernelRngTrain:
    """cfg.train hook for the oracle that rebuilds the kernels' keep masks from (seed, site, element index) with the
    tensor restatement of the counter RNG (pq3d_b200/rng.py) — element indexing as documented in include/pq3d_b200.h."""

    def __init__(self, enc, seed, p, B, N, H):
        from pq3d_b200 import rng
        self.rng, self.enc, self.seed, self.p, self.B, self.N, self.H = rng, enc, seed, p, B, N, H
        self.program = [g for g in enc._program() if len(g) > 0]
        self.keeps = list(enc.last_memory_keep)

    def _apply(self, x, keep):
        return x * keep.to(x.dtype) / (1.0 - self.p)

    def sublayer(self, layer, kind, x):
        rng, (B, N, D) = self.rng, x.shape
        R = B * N
        e = torch.arange(R * D, device=x.device).view(B, N, D)
        if isinstance(kind, tuple):
            gi = next(i for i, g in enumerate(self.program) if kind[1] in g)
            e = e + self.program[gi].index(kind[1]) * R * D
            site = rng.site(layer, rng.SITE_CA_SUBLAYER + gi)
        else:
            site = rng.site(layer, rng.SITE_SA_SUBLAYER if kind == "sa" else rng.SITE_FFN_SUBLAYER)
        return self._apply(x, rng.keep_mask(self.seed, site, e, self.p))

    def probs(self, layer, kind, P):
        rng, H = self.rng, self.H
        BH, L, S2 = P.shape
        S = S2 - 1 if isinstance(kind, tuple) else S2            # cross-attention carries the zero-attn column
        s_pad = (S + 127) // 128 * 128
        rows = torch.arange(BH * L, device=P.device).view(BH, L, 1)          # (b*H + h)*N + n
        e = rows * s_pad + torch.arange(S, device=P.device).view(1, 1, S)
        site = (rng.site(layer, rng.SITE_CA_PROBS + self.enc.memories.index(kind[1])) if isinstance(kind, tuple)
                else rng.site(layer, rng.SITE_SA_PROBS))
        keep = rng.keep_mask(self.seed, site, e, self.p)
        if S2 != S:
            keep = torch.cat([keep, torch.ones_like(keep[..., :1])], -1)
        return self._apply(P, keep)

    def hidden(self, layer, h):
        e = torch.arange(h.numel(), device=h.device).view(h.shape)
        return self._apply(h, self.rng.keep_mask(self.seed, self.rng.site(layer, self.rng.SITE_FFN_HIDDEN), e, self.p))

    def memory_keep(self, layer, memories, B):
        return self.keeps.pop(0)


def run_model_training_case(stage, device, ctx):
    """Query3DUnified in .train() (dropouts off): gradients of every model parameter vs autograd through the oracle's
    query3d_unified_forward, fp32, with the oracle under bf16 autocast as the yardstick.  `ctx`: a context manager
    active during the product's forward + backward (the CPU kernel emulation, or a null context on the GPU)."""
    from oracle import restatement as O
    from pq3d_b200 import synth
    import _cases as C
    from pq3d_b200.query3d_unified import Query3DUnified
    if stage == "stage1_mask":
        case = dict(base="c2", over=dict(B=2, N=16, S=72, num_layers=2, use_self_mask=True), dim_loc=3, heads=["mask"],
                    skip=False, wseed=41, sharp=1.0)
    else:
        case = dict(base="c3", over=dict(B=2, N=16, S=72, T=6, num_layers=2), dim_loc=6, heads=["ground"], skip=False,
                    wseed=42, sharp=1.0)
    w, cfg = C.model_cfg(case)
    for m in w.memories:
        if m != "prompt":
            cfg["model"][f"{m}_encoder"]["args"]["dropout"] = 0.0
    if "mask" in case["heads"]:
        cfg["model"]["mask_head"]["args"]["dropout"] = 0.0
    if "ground" in case["heads"]:
        cfg["model"]["ground_head"]["args"]["dropout"] = 0.0
    sd = C.to_dev(synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"]), device)
    model = Query3DUnified(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device).train()
    model.unified_encoder.train_dropout = 0.0
    if device == "cpu":
        model.unified_encoder.use_cuda_graph = False
        model.unified_encoder.train_streams = False
    d = C.to_dev(synth.make_model_data_dict(w, cfg), device)
    g = torch.Generator().manual_seed(9)

    def loss_of(out):
        tot = 0.0
        if "mask" in case["heads"]:
            for c, m in zip(out["predictions_class"], out["predictions_mask"]):
                uc = torch.randn(c.shape, generator=torch.Generator().manual_seed(c.shape[0] * 7 + 1)).to(device)
                um = torch.randn(m.shape, generator=torch.Generator().manual_seed(m.shape[1] * 3 + 2)).to(device)
                tot = tot + (c.float().masked_fill(~torch.isfinite(c.float()), 0.0) * uc).sum() * 0.1
                tot = tot + (m.float().clamp_min(-100.0) * um).sum() * 0.1
        if "ground" in case["heads"]:
            lg = out["ground_logits"].float()
            ug = torch.randn(lg.shape, generator=torch.Generator().manual_seed(5)).to(device)
            tot = tot + (lg.masked_fill(~torch.isfinite(lg), 0.0) * ug).sum()
        return tot
    with ctx:
        out = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()})
        loss_of(out).backward()
    def oracle(autocast):
        sdd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith("gauss_B") else v.clone())
               for k, v in sd.items()}
        with torch.autocast(device, dtype=torch.bfloat16, enabled=autocast):
            ref = O.query3d_unified_forward(sdd, C.oracle_model_cfg(w, cfg),
                                            {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()})
        loss_of(ref).backward()
        return {k: v.grad for k, v in sdd.items() if v.is_floating_point()}
    g32, g16 = oracle(False), oracle(True)
    l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()     # noqa: E731
    worst = []
    for k, p in model.named_parameters():
        r = g32[k]
        if r is None or float(r.norm()) == 0.0 or k.endswith("w_ks.bias"):
            continue
        assert p.grad is not None, f"{k}: no gradient"
        worst.append((l2(p.grad, r), l2(g16[k], r), k))
    worst.sort(reverse=True)
    print(worst[:5])
    assert len(worst) > 20
    for e, e16, k in worst:          # the reference's own bf16 path is the yardstick, as on the GPU
        assert e <= 1.5 * e16 + 3e-2, f"{k}: {e:.3e} (autocast oracle {e16:.3e})"


def run_prompt_loc_case(dim_loc, device, ctx):
    """Query3DUnified forward (eval) on a batch mixing text prompts (precomputed features) and location prompts (encoded
    by the coordinate encoders, model/query3d_unified.py:80-108) against the oracle's model forward."""
    from oracle import restatement as O
    from pq3d_b200 import synth
    import _cases as C
    from pq3d_b200.query3d_unified import Query3DUnified
    case = dict(base="c3", over=dict(B=4, N=16, S=72, T=6, num_layers=2), dim_loc=dim_loc, heads=["ground"], skip=False,
                wseed=44, sharp=1.0)
    w, cfg = C.model_cfg(case)
    sd = C.to_dev(synth.draw_state_dict(synth.model_param_shapes(cfg), case["wseed"], case["sharp"]), device)
    model = Query3DUnified(cfg).eval()
    model.load_state_dict(sd, strict=True)
    model = model.to(device)
    if device == "cpu":
        model.unified_encoder.use_cuda_graph = False
    d = C.to_dev(synth.make_model_data_dict(w, cfg), device)
    g = torch.Generator().manual_seed(4)
    d["prompt"] = (torch.rand(w.B, w.T, generator=g) * 3).to(device)
    d["prompt_type"] = torch.tensor([1, 3, 3, 1]).to(device)
    clone = lambda dd: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in dd.items()}     # noqa: E731
    d1, d2 = clone(d), clone(d)
    with ctx, torch.no_grad():
        out = model(d1)
    with torch.no_grad():
        ref = O.query3d_unified_forward(sd, C.oracle_model_cfg(w, cfg), d2)
        with torch.autocast(device, dtype=torch.bfloat16):
            ref16 = O.query3d_unified_forward(sd, C.oracle_model_cfg(w, cfg), clone(d))
    assert torch.equal(d1["prompt_pad_masks"], d2["prompt_pad_masks"])          # slot mask write-back (:102, :107)
    a, b, b16 = out["ground_logits"].float(), ref["ground_logits"].float(), ref16["ground_logits"].float()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin)
    e = ((a[fin] - b[fin]).abs().max() / b[fin].abs().max()).item()
    e16 = ((b16[fin] - b[fin]).abs().max() / b[fin].abs().max()).item()
    assert e <= 1.25 * e16 + 5e-3, (e, e16)
