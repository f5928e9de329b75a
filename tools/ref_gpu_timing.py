"""The "reference PyTorch GPU path" of SURVEY.md §8(d): the oracle restatement (== the reference modules to 1e-5) run
eagerly on the same B200 under torch.autocast(bf16) — what a user of the reference gets on this GPU — timed with CUDA
events on BASELINE config 3's per-GPU shard, next to nothing of ours.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import restatement as O
from pq3d_b200 import synth

w = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
dev = "cuda"
sd = {k: v.to(dev) for k, v in synth.decoder_state_dict(w, seed=0).items()}
cfg = O.DecoderCfg(**w.decoder_kwargs())
inp, pw, _ = synth.make_decoder_inputs(w, device=dev)
res = {}
for name, autocast, iters in (("autocast_bf16", True, 30), ("fp32", False, 10)):
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        for _ in range(5):
            O.query_mask_encoder(sd, cfg, synth.clone_input_dict(inp), pw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            O.query_mask_encoder(sd, cfg, synth.clone_input_dict(inp), pw)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[name] = {"ms_per_step": ms, "queries_per_s": w.B * w.N / (ms * 1e-3), "iters": iters}
print(json.dumps({"what": "oracle restatement of the reference decoder, eager PyTorch on this GPU", "workload": w.name,
                  "scenes": w.B, "queries": w.N, "seg_tokens": w.S, **res}))
