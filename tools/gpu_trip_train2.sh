#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?"
grep -E "forward:|passed|failed|losses|Error" gpurun_out/pytest_train.log | tail -20
timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 600 python tools/train_profile.py c5 > gpurun_out/train_profile.log 2>&1; echo "prof rc=$?"; head -60 gpurun_out/train_profile.log | cut -c1-200
