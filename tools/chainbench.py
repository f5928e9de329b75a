"""Effective per-kernel cost inside a CUDA graph of dependent launches (what the decoder body pays)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pq3d_b200 import ops

dev = "cuda"


def graph_time(fn, n=40, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n * 1e3


R, D, B, H, Nq = 400, 768, 4, 12, 100
x = torch.randn(R, 2048, device=dev).bfloat16()
for (N, K, bn, f32) in ((768, 768, 64, True), (768, 768, 128, True), (2304, 768, 64, False), (2304, 768, 128, False),
                        (2048, 768, 64, False), (2048, 768, 128, False), (768, 2048, 64, True), (1536, 768, 64, False)):
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.zeros(N, device=dev)
    C = torch.empty(R, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
    A = x[:, :K]
    us = graph_time(lambda: ops.linear(A, W, C, M=R, N=N, K=K, bias=b, block_n=bn))
    us2 = graph_time(lambda: ops.linear(A, W, C, M=R, N=N, K=K, bias=b, block_n=bn, w_const=True))
    print(f"linear M=400 N={N} K={K} bn={bn} {'f32' if f32 else 'bf16'}: {us:.2f} us/launch in-graph; W prefetched before the PDL wait: {us2:.2f}")
W = (torch.randn(3 * 768, 768, device=dev) * 0.02).bfloat16()
O = torch.randn(3 * R, 768, device=dev).bfloat16()
y = torch.empty(3, R, 768, device=dev)
bo = torch.zeros(3, 768, device=dev)
us = graph_time(lambda: ops.linear(O, W, y, M=R, N=768, K=768, bias=bo, bias_group_stride=768, groups=3, a_group_rows=R,
                                   w_group_rows=768, ldc=768, c_group_stride=R * 768))
print(f"grouped out-proj 3x[400x768x768]: {us:.2f} us")
yy, rr, pp = torch.randn(3, R, D, device=dev), torch.randn(R, D, device=dev), torch.randn(R, D, device=dev)
g_, b_ = torch.ones(3, D, device=dev), torch.zeros(3, D, device=dev)
o32, o16, op16 = torch.empty(R, D, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev)
for G in (1, 3):
    us = graph_time(lambda: ops.add_layernorm(yy, rr, g_, b_, 1e-5, R, D, G=G, y_group_stride=R * D, pos=pp, out_f32=o32, out_bf16=o16, out_pos_bf16=op16))
    print(f"add_layernorm G={G}: {us:.2f} us")
L = 4
for (S, nm, zero) in ((32, 1, True), (100, 1, False), (2048, 3, True), (2048, 1, True)):
    Sp = ops.pad8(S)
    Q = torch.randn(B * Nq, nm * D, device=dev).bfloat16()
    mems = []
    for i in range(nm):
        Kb = torch.randn(B * Sp, L * D, device=dev).bfloat16()
        Vt = torch.randn(L * D, B * Sp, device=dev).bfloat16()
        bits = ops.pack_mask(torch.rand(B, S, device=dev) < 0.1)
        mems.append(ops.AttnMemory(Kb, D, Vt, D, S, Sp, bits, bits.stride(0), 0, 0))
    Ot = torch.empty(nm, B * Nq, D, dtype=torch.bfloat16, device=dev)
    us = graph_time(lambda: ops.attention(Q, D, mems, Ot, B * Nq * D, B, H, Nq, zero), n=20)
    print(f"attention S={S} mems={nm}: {us:.2f} us")
c = torch.empty(R, D, dtype=torch.bfloat16, device=dev)
us = graph_time(lambda: ops.cast_bf16(rr, c))
print(f"cast_bf16 (tiny elementwise floor): {us:.2f} us")
