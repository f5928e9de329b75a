"""Where do the ~5 us of one query-side GEMM (M = 400, bn = 64) go?  Per-CTA cycle stamps (pq3d_debug_set_timeline):
kernel body start (after griddepcontrol.wait) -> first operands landed -> last MMA issued -> epilogue done -> exit,
for the last launch of a back-to-back sequence (warm instruction cache / L2) and for a lone launch after an L2 flush."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import microbench as mb

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, K, f32) in ((768, 768, True), (2304, 768, False), (768, 2048, True)):
    fn = mb.gemm_case(400, N, K, 64, out_fp32=f32)
    tiles = 4 * (N // 64)

    def warm(fn=fn):
        for _ in range(20):
            fn()
    warm()
    mb.timeline(fn, tiles, f"M=400 N={N} K={K} warm (after 20 launches)")

    def cold(fn=fn):
        flush.zero_()
        fn()
    mb.timeline(cold, tiles, f"M=400 N={N} K={K} after an L2 flush")
