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
elif which == "attn_bwd":
    # one scene-memory cross-attention backward at the config-5 shard shape (2 scenes, S = 2048)
    Bt = 2
    Q = torch.randn(Bt * Nq, D, device=dev).bfloat16()
    Kb = torch.randn(Bt * S, L * D, device=dev).bfloat16()
    Vb = torch.randn(Bt * S, L * D, device=dev).bfloat16()
    Vt = Vb[:, :D].t().contiguous()
    bits = ops.pack_mask(torch.rand(Bt, S, device=dev) < 0.1)
    O = torch.empty(1, Bt * Nq, D, dtype=torch.bfloat16, device=dev)
    st_m, st_l = torch.empty(Bt, H, Nq, device=dev), torch.empty(Bt, H, Nq, device=dev)
    ops.attention(Q, 0, [ops.AttnMemory(Kb, 0, Vt, 0, S, S, bits, bits.stride(0), 0, 0)], O, Bt * Nq * D, Bt, H, Nq, True,
                  stats=(st_m, st_l))
    dO = torch.randn(Bt * Nq, D, device=dev).bfloat16()
    delta = torch.empty(Bt, H, Nq, device=dev)
    ops.attn_delta(dO, O[0], delta, Bt, H, Nq)
    dK, dV = torch.zeros_like(Kb), torch.zeros_like(Vb)
    dQ = torch.zeros(Bt * Nq, D, device=dev)
    for _ in range(5):
        ops.attention_bwd(Q, 0, dO, 0, Kb, 0, Vb, 0, S, S, st_m, st_l, delta, dK, 0, dV, 0, dQ, 0, Bt, H, Nq, mask_bits=bits,
                          mask_strides=(bits.stride(0), 0, 0))
    torch.cuda.synchronize()
    import time
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(50):
        ops.attention_bwd(Q, 0, dO, 0, Kb, 0, Vb, 0, S, S, st_m, st_l, delta, dK, 0, dV, 0, dQ, 0, Bt, H, Nq, mask_bits=bits,
                          mask_strides=(bits.stride(0), 0, 0))
    t1.record(); torch.cuda.synchronize()
    print(f"attention_bwd alone: {t0.elapsed_time(t1) / 50 * 1e3:.1f} us per launch (B={Bt}, H={H}, N={Nq}, S={S})")
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
