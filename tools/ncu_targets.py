"""Launches exactly the two hot kernels at the bench shapes, for `ncu --set full` captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pq3d_b200 import ops

dev = "cuda"
B, S, D, L, H, Nq, nf = 4, 2048, 768, 4, 12, 100, 3
which = sys.argv[1]
if which == "gemm":
    M, Nn = B * S, L * D
    A = torch.randn(nf * M, D, device=dev).bfloat16()
    W = (torch.randn(nf * L * D, D, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(nf * L * D, device=dev)
    C = torch.empty(nf, M, Nn, dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        ops.linear(A, W, C, M=M, N=Nn, K=D, bias=bias, bias_group_stride=L * D, groups=nf, a_group_rows=M,
                   w_group_rows=L * D, ldc=Nn, c_group_stride=M * Nn)
else:
    Q = torch.randn(B * Nq, nf * D, device=dev).bfloat16()
    mems = []
    for i in range(nf):
        Kb = torch.randn(B * S, L * D, device=dev).bfloat16()
        Vt = torch.randn(L * D, B * S, device=dev).bfloat16()
        bits = ops.pack_mask(torch.rand(B, S, device=dev) < 0.1)
        mems.append(ops.AttnMemory(Kb, D, Vt, D, S, S, bits, bits.stride(0), 0, 0))
    O = torch.empty(nf, B * Nq, D, dtype=torch.bfloat16, device=dev)
    for _ in range(5):
        ops.attention(Q, D, mems, O, B * Nq * D, B, H, Nq, True)
torch.cuda.synchronize()
