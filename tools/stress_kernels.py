"""Which kernel stalls under multi-stream concurrency?  Per stream one CUDA graph of n launches of ONE kernel kind (own
buffers), replayed round-robin over NS streams for SECONDS; a monitor thread reports a stall (no replay returning for 8 s)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pq3d_b200 import ops
dev = torch.device("cuda", 0)
kind = sys.argv[1]
ns = int(os.environ.get("NS", "4"))
secs = float(os.environ.get("SECONDS", "25"))
R, D, B, H, Nq, S, L = 400, 768, 4, 12, 100, 2048, 4


def make(kind):
    if kind == "gemm_pair":
        A = torch.randn(3 * B * S, D, device=dev).bfloat16()
        W = (torch.randn(3 * L * D, D, device=dev) * 0.02).bfloat16()
        b = torch.zeros(3 * L * D, device=dev)
        C = torch.empty(3, B * S, L * D, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.linear(A, W, C, M=B * S, N=L * D, K=D, bias=b, bias_group_stride=L * D, groups=3, a_group_rows=B * S,
                                   w_group_rows=L * D, ldc=L * D, c_group_stride=B * S * L * D)), 2
    if kind == "gemm64":
        A = torch.randn(R, D, device=dev).bfloat16()
        W = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
        b = torch.zeros(D, device=dev)
        C = torch.empty(R, D, dtype=torch.float32, device=dev)
        return (lambda: ops.linear(A, W, C, M=R, N=D, K=D, bias=b, w_const=True)), 40
    if kind == "attention":
        Q = torch.randn(R, 3 * D, device=dev).bfloat16()
        mems = []
        for _ in range(3):
            Kb = torch.randn(B * S, L * D, device=dev).bfloat16()
            Vt = torch.randn(L * D, B * S, device=dev).bfloat16()
            bits = ops.pack_mask(torch.rand(B, S, device=dev) < 0.1)
            mems.append(ops.AttnMemory(Kb, D, Vt, D, S, S, bits, bits.stride(0), 0, 0))
        O = torch.empty(3, R, D, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.attention(Q, D, mems, O, R * D, B, H, Nq, True)), 12
    if kind == "attention_small":
        Q = torch.randn(R, 2 * D, device=dev).bfloat16()
        Np = ops.pad8(Nq)
        Vt = torch.randn(D, B * Np, device=dev).bfloat16()
        qb = ops.pack_mask(torch.zeros(B, Nq, dtype=torch.bool, device=dev))
        mem = ops.AttnMemory(Q, D, Vt, 0, Nq, Nq, qb, qb.stride(0), 0, 0, Vt_pitch=Np)
        O = torch.empty(1, R, D, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.attention(Q, 0, [mem], O, R * D, B, H, Nq, False)), 20
    if kind == "ln":
        y, r_, p_ = torch.randn(3, R, D, device=dev), torch.randn(R, D, device=dev), torch.randn(R, D, device=dev)
        g_, b_ = torch.ones(3, D, device=dev), torch.zeros(3, D, device=dev)
        o32, o16, op16 = torch.empty(R, D, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.add_layernorm(y, r_, g_, b_, 1e-5, R, D, G=3, y_group_stride=R * D, pos=p_, out_f32=o32, out_bf16=o16, out_pos_bf16=op16)), 40
    if kind == "mix":            # chain-like: gemm64 -> attention_small -> gemm64 -> ln
        f1, _ = make("gemm64"); f2, _ = make("attention_small"); f3, _ = make("ln"); f4, _ = make("attention")
        def f():
            f1(); f4(); f1(); f3(); f1(); f2(); f1(); f3()
        return f, 5
    raise KeyError(kind)


streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
graphs, keep = [], []
for st in streams:
    fn, n = make(kind)
    keep.append(fn)
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
    graphs.append(g)
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
            print(f"STALL {kind} at replay {last}: gpu util {u.gpu}% power {pynvml.nvmlDeviceGetPowerUsage(h)/1000:.0f} W; "
                  f"streams done: {[s.query() for s in streams]}", flush=True)
            os._exit(3)
        last = progress["i"]


threading.Thread(target=monitor, daemon=True).start()
t0 = time.time()
i = 0
while time.time() - t0 < secs:
    for _ in range(64):
        with torch.cuda.stream(streams[i % ns]):
            graphs[i % ns].replay()
        i += 1
        progress["i"] = i; progress["t"] = time.time()
    if i % 4096 == 0:
        torch.cuda.synchronize()
torch.cuda.synchronize()
print(f"OK {kind}: {i} replays x {n} launches in {time.time()-t0:.1f} s", flush=True)
