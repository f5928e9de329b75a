"""Loss trajectories of the training step: oracle autograd vs ours (eager, several AdamW flavours) vs graphed."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import restatement as O
from pq3d_b200 import synth
from pq3d_b200.query_encoder import QueryMaskEncoder
from pq3d_b200.training import GraphedTrainStep

DEV = "cuda"
w = synth.Workload("tgraph", 2, 100, 256, ["mv", "pc", "voxel", "prompt"], "mixed", T=16, num_layers=2)
sd = synth.decoder_state_dict(w, seed=5)
inp, pw, _ = synth.make_decoder_inputs(w, device=DEV)
g = torch.Generator().manual_seed(12)
q, qm, qp = inp["query"]
inp["query"] = (torch.randn(q.shape, generator=g).to(DEV) * 0.5, qm, qp)
target = torch.randn(q.shape, generator=g).to(DEV)
loss_fn = lambda out, tgt: ((out - tgt) ** 2).mean()
STEPS, LR = 7, 1e-3

sdd = {k: v.to(DEV).clone().requires_grad_(True) for k, v in sd.items()}
opt = torch.optim.AdamW(list(sdd.values()), lr=LR, betas=(0.9, 0.98))
cfg = O.DecoderCfg(**w.decoder_kwargs())
tr = []
for _ in range(STEPS):
    opt.zero_grad(set_to_none=True)
    loss = loss_fn(O.query_mask_encoder(sdd, cfg, synth.clone_input_dict(inp), pw)[0], target)
    loss.backward(); opt.step(); tr.append(round(loss.item(), 4))
print("oracle fp32 autograd + AdamW      ", tr)

for name, kw in (("ours eager, foreach AdamW", {}), ("ours eager, fused", dict(fused=True)),
                 ("ours eager, fused+capturable", dict(fused=True, capturable=True)), ("ours graphed", dict(fused=True, capturable=True))):
    enc = QueryMaskEncoder(None, **w.decoder_kwargs())
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(DEV).train(); enc.train_dropout = 0.0
    opt = torch.optim.AdamW(enc.parameters(), lr=LR, betas=(0.9, 0.98), **kw)
    tr = []
    if name == "ours graphed":
        step = GraphedTrainStep(enc, opt, loss_fn, warmup=2)
        for _ in range(STEPS):
            tr.append(round(step(inp, pw, target).item(), 4))
    else:
        v0 = [p._version for p in enc.parameters()][:3]
        for _ in range(STEPS):
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(enc(synth.clone_input_dict(inp), pw)[0], target)
            loss.backward(); opt.step(); tr.append(round(loss.item(), 4))
        print("   versions", v0, "->", [p._version for p in enc.parameters()][:3])
    print(f"{name:34s}", tr)
