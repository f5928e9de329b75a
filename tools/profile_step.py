"""Per-kernel GPU time of one decoder step (torch.profiler / CUPTI, real clocks, no serialisation).
Run eagerly (no CUDA graph) so every kernel is a separate launch record."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from pq3d_b200 import synth
from pq3d_b200.query_encoder import QueryMaskEncoder

w = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c3")
enc = QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
enc.load_state_dict(synth.decoder_state_dict(w, seed=0), strict=True)
enc = enc.cuda()
enc.use_cuda_graph = False
inp, pw, _ = synth.make_decoder_inputs(w, device="cuda")
with torch.no_grad():
    for _ in range(3):
        enc(synth.clone_input_dict(inp), pw)
    torch.cuda.synchronize()
    steps = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            enc(synth.clone_input_dict(inp), pw)
        torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.split("(")[0][:70]
        tot[name] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        cnt[name] += 1
T = sum(tot.values())
print(f"GPU kernel time per step: {T / steps:.1f} us over {sum(cnt.values()) // steps} kernels")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"  {v / steps:9.1f} us/step {100 * v / T:5.1f}%  n/step={cnt[k] / steps:5.1f}  avg={v / cnt[k]:7.2f} us  {k}")
