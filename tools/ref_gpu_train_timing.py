"""The reference PyTorch GPU path for the TRAINING step of BASELINE config 5's shard: eager autograd through the oracle
restatement under torch.autocast(bf16) + fused AdamW on the same B200, dropout 0.1 at the reference's sites."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import restatement as O
from pq3d_b200 import synth

w = synth.workload("c5")
dev = "cuda"
sd = {k: v.to(dev).requires_grad_(True) for k, v in synth.decoder_state_dict(w, seed=0).items()}
cfg = O.DecoderCfg(**w.decoder_kwargs())


class Drop:
    sublayer = staticmethod(lambda layer, kind, x: F.dropout(x, 0.1, True))
    probs = staticmethod(lambda layer, kind, x: F.dropout(x, 0.1, True))
    hidden = staticmethod(lambda layer, x: F.dropout(x, 0.1, True))
    memory_keep = staticmethod(lambda layer, memories, B: None)


cfg.train = Drop
inp, pw, _ = synth.make_decoder_inputs(w, device=dev)
target = torch.randn(w.B, w.N, w.hidden_size, device=dev)
opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, betas=(0.9, 0.98), fused=True)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.query_mask_encoder(sd, cfg, synth.clone_input_dict(inp), pw)[0]
    ((out.float() - target) ** 2).mean().backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(15):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 15
print(json.dumps({"what": "oracle restatement, eager autograd under autocast(bf16) + fused AdamW on this GPU, dropout 0.1",
                  "workload": "c5", "scenes": w.B, "ms_per_step": ms, "queries_per_s": w.B * w.N / (ms * 1e-3)}))
