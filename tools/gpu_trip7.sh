#!/bin/bash
mkdir -p gpurun_out
bash tests/run_kernel_checks.sh > /dev/null 2>&1; cat gpurun_out/kernel_checks.summary | grep -v "exit 0"
grep -E "Error|error|assert|timed out|ragged tiles" gpurun_out/kernel_checks.log | head -20
for cl in 1 0; do echo "== PQ3D_GEMM_CLUSTER=$cl"; PQ3D_GEMM_CLUSTER=$cl timeout 600 python tools/microbench.py 2>&1 | grep -E "bn=256 K=|timeline 8192|last MMA|cycles per tile|exit  " | head -12; done
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -5
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['serial_value']),'roofline',round(d['roofline']['achieved']),round(d['roofline']['frac'],3), d['roofline']['kernel'])"; tail -3 gpurun_out/bench.err
