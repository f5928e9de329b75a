#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "err\(|passed|failed|invariance|equivariance|attn-mask bits" gpurun_out/pytest_gpu.log | tail -40
cat gpurun_out/microbench.log
