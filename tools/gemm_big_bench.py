"""The hoisted K / V^T projections of config 3 (3 scene memories x [8192 x 3072 x 768], grouped) timed inside a CUDA
graph, with and without CTA pairs.  Run once per PQ3D_GEMM_EPI_WARPS setting (the switch is read once per process)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pq3d_b200 import ops

dev = "cuda"
nf, BS, D, L = 3, 8192, 768, 4
bf16 = torch.bfloat16
x = (torch.randn(nf * BS, D, device=dev) * 0.5).to(bf16)
wk = (torch.randn(nf * L * D, D, device=dev) * 0.03).to(bf16)
bk = torch.randn(nf * L * D, device=dev)
K_all = torch.empty(nf, BS, L * D, dtype=bf16, device=dev)
Vt_all = torch.empty(nf, L * D, BS, dtype=bf16, device=dev)


def k_proj(no_pairs):
    ops.linear(x, wk, K_all, M=BS, N=L * D, K=D, bias=bk, bias_group_stride=L * D, groups=nf, a_group_rows=BS,
               w_group_rows=L * D, ldc=L * D, c_group_stride=BS * L * D, no_pairs=no_pairs)


def vt_proj(no_pairs):
    ops.linear(wk, x, Vt_all, M=L * D, N=BS, K=D, bias=bk, bias_along_m=True, bias_group_stride=L * D, groups=nf,
               a_group_rows=L * D, w_group_rows=BS, ldc=BS, c_group_stride=L * D * BS, no_pairs=no_pairs)


def graph_time(fn, n=6, reps=6):
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


flops = 2.0 * nf * BS * L * D * D
mode = os.environ.get("PQ3D_GEMM_EPI_WARPS", "default")
ref = (x.float().view(nf, BS, D) @ wk.float().view(nf, L * D, D).transpose(1, 2) + bk.view(nf, 1, L * D))
for no_pairs in (False, True):
    for name, fn, out, want in (("K", k_proj, K_all, ref), ("V^T", vt_proj, Vt_all, ref.transpose(1, 2))):
        out.zero_()
        us = graph_time(lambda: fn(no_pairs))
        err = ((out.float() - want).abs().max() / want.abs().max()).item()
        print(f"epi_warps={mode} {'single CTA' if no_pairs else 'CTA pairs '} {name:>3} projection: {us:7.2f} us  "
              f"{flops / us / 1e6:7.1f} TFLOP/s   rel err {err:.2e}")
