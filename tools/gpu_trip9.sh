#!/bin/bash
mkdir -p gpurun_out
for wl in c4 c2 c1; do
timeout 900 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_$wl.log').read().strip().splitlines()[-1]);print('$wl value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['serial_value']),'launches/step',d['gpu_launches']/d['steps'], d['config']['workload'])"; tail -3 gpurun_out/bench_$wl.err
done
