#!/bin/bash
for f in 0 1 2; do echo "== PQ3D_GEMM_DEBUG=$f"; PQ3D_GEMM_DEBUG=$f timeout 600 python tools/microbench.py 2>&1 | grep -E "bn=256 K=  768|bn=256 K= 2048|timeline 8192|last MMA|cycles per tile|exit  " | head -8; done
timeout 100 python tests/kernel_checks.py attn_ragged_tiles 2>&1 | tail -3
