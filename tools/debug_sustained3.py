"""Long 4-stream replay loop with a monitor thread: when no step completes for 8 s, print GPU utilisation / power (NVML)
and which streams still have work, then exit."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pq3d_b200 import synth
from pq3d_b200.query_encoder import QueryMaskEncoder
dev = torch.device("cuda", 0)
w = synth.workload("c3")
enc = QueryMaskEncoder(None, **w.decoder_kwargs()).eval()
enc.load_state_dict(synth.decoder_state_dict(w, seed=0), strict=True)
enc = enc.to(dev)
inp_host, pw_host, _ = synth.make_decoder_inputs(w)
memo = {}
def g(t):
    if isinstance(t, torch.Tensor):
        if id(t) not in memo: memo[id(t)] = t.to(dev)
        return memo[id(t)]
    if isinstance(t, (list, tuple)): return type(t)(g(x) for x in t)
    return t
inp = {k: g(v) for k, v in inp_host.items()}
pw = pw_host.to(dev)
ns = int(os.environ.get("NS", "4"))
cur = torch.cuda.current_stream()
streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
mode = os.environ.get("MODE", "full")          # full = whole-forward graph, body = eager prologue + body graph, eager
def step():
    with torch.no_grad():
        d = synth.clone_input_dict(inp)
        if mode == "body":
            d["query"] = tuple(t.clone() for t in d["query"])     # fresh tensors: no whole-forward graph
        return enc(d, pw)[0]
TRACE = {}
if mode == "eager":
    enc.use_cuda_graph = False
    from pq3d_b200 import ops as _ops
    import collections
    for _n in ["linear", "attention", "spatial_bias", "ingest_memory", "ingest_memories", "add_layernorm", "pack_mask", "cast_bf16"]:
        _f = getattr(_ops, _n)
        def _w(*a, _f=_f, _n=_n, **k):
            r = _f(*a, **k)
            st = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(st)
            info = _n + (f" M={k.get('M')} N={k.get('N')} K={k.get('K')} g={k.get('groups', 1)}" if _n == "linear" else "")
            TRACE.setdefault(st.cuda_stream, collections.deque(maxlen=120)).append((info, ev))
            return r
        setattr(_ops, _n, _w)
for st in streams:
    st.wait_stream(cur)
    with torch.cuda.stream(st):
        for _ in range(5): step()
torch.cuda.synchronize()
progress = {"i": 0, "t": time.time()}
def monitor():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    last = -1
    while True:
        time.sleep(2)
        if progress["i"] == last and time.time() - progress["t"] > 8:
            u = pynvml.nvmlDeviceGetUtilizationRates(h)
            print(f"STALL at step {last}: gpu util {u.gpu}% mem util {u.memory}% power {pynvml.nvmlDeviceGetPowerUsage(h)/1000:.0f} W "
                  f"sm clock {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz", flush=True)
            for h_, dq in TRACE.items():
                pend = [n for n, e in dq if not e.query()]
                if pend:
                    print(f"stream {hex(h_)}: first pending ops: {pend[:4]} ({len(pend)} pending)", flush=True)
            try:
                print("stream.query():", [s.query() for s in streams], flush=True)
            except Exception as e:
                print("query error", repr(e), flush=True)
            os._exit(3)
        last = progress["i"]
threading.Thread(target=monitor, daemon=True).start()
N = int(os.environ.get("STEPS", "60000"))
sync_every = int(os.environ.get("SYNC", "0"))
t0 = time.perf_counter()
for i in range(N):
    with torch.cuda.stream(streams[i % ns]):
        step()
    progress["i"] = i; progress["t"] = time.time()
    if sync_every and (i + 1) % sync_every == 0:
        torch.cuda.synchronize()
    if (i + 1) % 10000 == 0:
        torch.cuda.synchronize()
        print(f"{i+1} steps {(time.perf_counter()-t0)/(i+1)*1e3:.3f} ms/step", flush=True)
torch.cuda.synchronize()
print("DONE", flush=True)
