#!/bin/bash
mkdir -p gpurun_out
bash tests/run_kernel_checks.sh > /dev/null 2>&1; cat gpurun_out/kernel_checks.summary
grep -E "Error|error|assert|timed out" gpurun_out/kernel_checks.log | head -20
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"
grep -v "^  N=\|empty launch\|add_layernorm G\|bn=64: 148\|bn=128: 148" gpurun_out/microbench.log
timeout 600 python tools/profile_step.py c3 > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"; cat gpurun_out/profile_step.log | tail -25
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -5
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
