"""Where does a step of the model-boundary e2e loop go?  host enqueue time of Query3DUnified.forward, its device
time, and the H2D copy alone (pinned -> device), on config 3.  Prints one line each."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pq3d_b200 import synth
from pq3d_b200.query3d_unified import Query3DUnified

dev = torch.device("cuda", 0)
w = synth.workload(sys.argv[1] if len(sys.argv) > 1 else "c3", None)
mcfg = synth.model_cfg_dict(w, dim_loc=3, heads=("ground",))
model = Query3DUnified(mcfg).eval()
msd = synth.draw_state_dict(synth.model_param_shapes(mcfg), 0)
msd.update({"unified_encoder." + k: v for k, v in synth.decoder_state_dict(w, seed=0).items()})
model.load_state_dict(msd, strict=True)
model = model.to(dev)
dd_host = synth.make_model_data_dict(w, mcfg, rank=0)
dd_pin = {k: v.pin_memory() for k, v in dd_host.items() if isinstance(v, torch.Tensor)}
nbytes = sum(t.numel() * t.element_size() for t in dd_pin.values())
for k, v in dd_pin.items():
    print(f"  {k}: {tuple(v.shape)} {v.dtype} {v.numel() * v.element_size() / 1e6:.2f} MB")
dd = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in dd_pin.items()}


def copy_in():
    for k, v in dd_pin.items():
        dd[k].copy_(v, non_blocking=True)


def ev_time(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


copy_in()
ms, host = ev_time(copy_in, 20)
print(f"H2D {nbytes / 1e6:.1f} MB in {len(dd_pin)} copies: {ms:.3f} ms device ({nbytes / ms / 1e6:.1f} GB/s), host enqueue {host:.3f} ms")
for graph in (False, True):
    model.use_cuda_graph = graph
    with torch.no_grad():
        for _ in range(6):
            model(dict(dd))["ground_logits"]
        ms, host = ev_time(lambda: model(dict(dd))["ground_logits"], 30)
    print(f"Query3DUnified.forward (whole-model graph {graph}): {ms:.3f} ms device per forward, host enqueue {host:.3f} ms")
from pq3d_b200 import ops
n0 = ops.LAUNCHES
with torch.no_grad():
    model(dict(dd))
print("launches per forward (our kernels, graph replays count their captured launches):", ops.LAUNCHES - n0)
