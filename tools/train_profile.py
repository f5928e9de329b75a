"""Per-kernel time of one training step (config 5 shard) via torch.profiler (kineto/CUPTI): where the step goes."""
import sys
import torch
sys.path.insert(0, ".")
from pq3d_b200 import synth
from pq3d_b200.query_encoder import QueryMaskEncoder

w = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c5")
dev = "cuda"
enc = QueryMaskEncoder(None, **w.decoder_kwargs())
enc.load_state_dict(synth.decoder_state_dict(w, seed=0), strict=True)
enc = enc.to(dev).train()
opt = torch.optim.AdamW(enc.parameters(), lr=1e-4, fused=True)
inp, pw, _ = synth.make_decoder_inputs(w, device=dev)
target = torch.randn(w.B, w.N, w.hidden_size, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    out = enc(synth.clone_input_dict(inp), pw)[0]
    loss = ((out - target) ** 2).mean()
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
opt.zero_grad(set_to_none=True)
ev[0].record()
out = enc(synth.clone_input_dict(inp), pw)[0]
loss = ((out - target) ** 2).mean()
ev[1].record()
loss.backward()
ev[2].record()
opt.step()
ev[3].record()
torch.cuda.synchronize()
print(f"fwd {ev[0].elapsed_time(ev[1]):.3f} ms  bwd {ev[1].elapsed_time(ev[2]):.3f} ms  adamw {ev[2].elapsed_time(ev[3]):.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
