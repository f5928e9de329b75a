"""GPU microbenchmarks + per-CTA timelines for the kernels (run under gpurun; prints to stdout)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pq3d_b200 import _lib, ops

dev = "cuda"


def timeit(fn, reps=30, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush is None:
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def gemm_case(M, N, K, bn, out_fp32=False, bias=True, groups=1):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.zeros(N, device=dev) if bias else None
    C = torch.empty(M, N, dtype=torch.float32 if out_fp32 else torch.bfloat16, device=dev)
    return lambda: ops.linear(A, W, C, M=M, N=N, K=K, bias=b, block_n=bn)


def timeline(fn, n_tiles, label):
    n_cta = min(n_tiles, 148)
    buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
    lib = _lib.lib()
    lib.pq3d_debug_set_timeline.argtypes = [ctypes.c_void_p]
    fn()
    torch.cuda.synchronize()
    lib.pq3d_debug_set_timeline(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.pq3d_debug_set_timeline(None)
    t = buf.view(n_cta, 8).cpu().double()
    d = lambda i: (t[:, i] - t[:, 1])  # noqa: E731
    print(f"[timeline {label}] persistent CTAs={n_cta}, tiles/CTA {t[:,4].min().item():.0f}..{t[:,4].max().item():.0f}; per-CTA cycles (median / max):")
    for name, i in (("first operands landed", 2), ("last MMA issued", 3), ("epilogue done", 5), ("exit", 6)):
        v = d(i)
        print(f"    {name:24s} {v.median().item():9.0f} / {v.max().item():9.0f}")
    print(f"    cycles per tile (exit / tiles): median {(d(6) / t[:,4]).median().item():.0f}")


if __name__ == "__main__":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print("== GEMM M=8192 N=3072 (K/V projection shape): us, TFLOP/s")
    for bn in (256, 128):
        for K in (64, 256, 768, 2048):
            us = timeit(gemm_case(8192, 3072, K, bn))
            print(f"  bn={bn} K={K:5d}: {us:8.1f} us  {2*8192*3072*K/us/1e6:7.1f} TF/s")
    us = timeit(gemm_case(8192, 3072, 768, 256), flush=flush)
    print(f"  bn=256 K=768 cold L2: {us:.1f} us")
    us = timeit(gemm_case(8192, 3072, 768, 256, bias=False))
    print(f"  bn=256 K=768 no bias: {us:.1f} us")
    us = timeit(gemm_case(8192, 3072, 768, 256, out_fp32=True))
    print(f"  bn=256 K=768 fp32 out: {us:.1f} us")
    print("== one wave: M=128*148")
    for bn in (256, 128, 64):
        us = timeit(gemm_case(128 * 148, bn, 768, bn))
        print(f"  bn={bn}: 148 CTAs, K=768: {us:.1f} us")
        us = timeit(gemm_case(128 * 148, bn, 64, bn))
        print(f"  bn={bn}: 148 CTAs, K=64: {us:.1f} us")
    print("== skinny (query side) M=400")
    for (N, K, bn) in ((768, 768, 64), (2304, 768, 64), (2048, 768, 64), (768, 2048, 64), (1536, 768, 64), (768, 768, 128)):
        us = timeit(gemm_case(400, N, K, bn, out_fp32=True))
        print(f"  N={N} K={K} bn={bn}: {us:.1f} us")
    timeline(gemm_case(8192, 3072, 768, 256), 12 * 64, "8192x3072x768 bn256")
    timeline(gemm_case(128 * 148, 256, 768, 256), 148, "one wave bn256 K=768")
    timeline(gemm_case(400, 2304, 768, 64, out_fp32=True), 36 * 4, "skinny 400x2304x768 bn64")

    print("== attention (c3 shapes)")
    B, H, Nq, D, L = 4, 12, 100, 768, 4
    for S in (128, 512, 2048, 4096):
        Sp = ops.pad8(S)
        Q = torch.randn(B * Nq, 3 * D, device=dev).bfloat16()
        mems = []
        for i in range(3):
            Kb = torch.randn(B * Sp, L * D, device=dev).bfloat16()
            Vt = torch.randn(L * D, B * Sp, device=dev).bfloat16()
            bits = ops.pack_mask(torch.rand(B, S, device=dev) < 0.1)
            mems.append(ops.AttnMemory(Kb, D, Vt, D, S, Sp, bits, bits.stride(0), 0, 0))
        O = torch.empty(3, B * Nq, D, dtype=torch.bfloat16, device=dev)
        us3 = timeit(lambda: ops.attention(Q, D, mems, O, B * Nq * D, B, H, Nq, True))
        us1 = timeit(lambda: ops.attention(Q, D, mems[:1], O, B * Nq * D, B, H, Nq, True))
        print(f"  S={S}: 3 memories {us3:.1f} us, 1 memory {us1:.1f} us")
        if S in (128, 2048):
            n_cta = H * B * 3
            tb = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
            lib = _lib.lib()
            lib.pq3d_debug_set_attention_timeline(tb.data_ptr())
            ops.attention(Q, D, mems, O, B * Nq * D, B, H, Nq, True)
            torch.cuda.synchronize()
            lib.pq3d_debug_set_attention_timeline(None)
            t = tb.view(n_cta, 8).cpu().double()
            for name, i in (("pass 1 done", 2), ("pass 2 done", 3), ("output stored", 4), ("exit", 5)):
                v = t[:, i] - t[:, 1]
                print(f"      [attention timeline S={S}] {name:14s} median {v.median().item():8.0f} max {v.max().item():8.0f} cycles after setup")
            span = (t[:, 0].max() - t[:, 0].min()) / 1e3
            print(f"      CTA start spread {span:.1f} us")
    print("== elementwise")
    R = 400
    y, res, pos = torch.randn(3, R, D, device=dev), torch.randn(R, D, device=dev), torch.randn(R, D, device=dev)
    g, b_ = torch.ones(3, D, device=dev), torch.zeros(3, D, device=dev)
    o32, o16, op16 = torch.empty(R, D, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev), torch.empty(R, D, dtype=torch.bfloat16, device=dev)
    print(f"  add_layernorm G=3: {timeit(lambda: ops.add_layernorm(y, res, g, b_, 1e-5, R, D, G=3, y_group_stride=R*D, pos=pos, out_f32=o32, out_bf16=o16, out_pos_bf16=op16)):.1f} us")
    print(f"  add_layernorm G=1: {timeit(lambda: ops.add_layernorm(y, res, g, b_, 1e-5, R, D, G=1, pos=pos, out_f32=o32, out_bf16=o16, out_pos_bf16=op16)):.1f} us")
    feat, p2 = torch.randn(4, 2048, D, device=dev), torch.randn(4, 2048, D, device=dev)
    xk, xv = torch.empty(4 * 2048, D, dtype=torch.bfloat16, device=dev), torch.empty(4 * 2048, D, dtype=torch.bfloat16, device=dev)
    us = timeit(lambda: ops.ingest_memory(feat, p2, xk, xv, 2048))
    print(f"  ingest 4x2048x768: {us:.1f} us = {(2*4*2048*768*4 + 2*4*2048*768*2)/us/1e3:.0f} GB/s")
    x = torch.empty(1, device=dev)
    print(f"  empty launch floor (cast 4 elems): {timeit(lambda: ops.cast_bf16(torch.zeros(4, device=dev), torch.empty(4, dtype=torch.bfloat16, device=dev))):.1f} us")
