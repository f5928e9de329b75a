#!/bin/bash
mkdir -p gpurun_out
for c in bwd_elementwise; do echo "=== $c"; timeout 200 python tests/kernel_checks.py $c 2>&1 | tail -5; done
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"
grep -vE "^\s*$" gpurun_out/pytest_train.log | tail -60
