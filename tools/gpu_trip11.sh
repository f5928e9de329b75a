#!/bin/bash
mkdir -p gpurun_out
for c in bgemm bwd_elementwise layernorm_bwd attention_bwd attn_basic gemm_shapes; do echo "=== $c"; timeout 200 python tests/kernel_checks.py $c 2>&1 | tail -12; done
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -5
