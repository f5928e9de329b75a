#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.log | tail -8
grep -E "mh_selfmask|model_stage" gpurun_out/pytest_gpu.log | head
for wl in c4; do
timeout 900 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_$wl.log').read().strip().splitlines()[-1]);print('$wl value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['serial_value']),'launches/step',d['gpu_launches']/d['steps'], d['config']['workload'])"; tail -3 gpurun_out/bench_$wl.err
done
