#!/bin/bash
mkdir -p gpurun_out
bash tests/run_kernel_checks.sh > /dev/null 2>&1; cat gpurun_out/kernel_checks.summary | grep -v "exit 0"
grep -E "Error|error|assert|timed out" gpurun_out/kernel_checks.log | head -20
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -5
timeout 600 python tools/profile_step.py c3 > gpurun_out/profile_step.log 2>&1; echo "profile rc=$?"; tail -14 gpurun_out/profile_step.log
for pdl in 1 0; do
PQ3D_PDL=$pdl timeout 900 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.log 2> gpurun_out/bench.err; echo "bench pdl=$pdl rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_pdl$pdl.log').read().strip().splitlines()[-1]);print('PDL=$pdl value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'roofline',round(d['roofline']['achieved']),d['roofline']['frac'], d['roofline']['kernel'])"; tail -3 gpurun_out/bench.err
done
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; grep -A8 "== attention" gpurun_out/microbench.log | head -30
