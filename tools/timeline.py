"""Kernel timeline of ONE decoder step replayed from its CUDA graph (torch.profiler / CUPTI): start offset, duration,
stream and name of every kernel, to see what actually overlaps.  PQ3D_PDL=0 gives honest per-kernel durations."""
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
inp, pw, _ = synth.make_decoder_inputs(w, device="cuda")
with torch.no_grad():
    for _ in range(6):
        enc(synth.clone_input_dict(inp), pw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        enc(synth.clone_input_dict(inp), pw)
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
end = max(e.time_range.end for e in ev)
print(f"{len(ev)} device records, span {end - t0:.1f} us")
for e in ev:
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f}  {e.name.split('(')[0][:60]}")
