#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list (+ optional full capture).  Logs -> gpurun_out/.
mkdir -p gpurun_out
echo "== pytest -m gpu" 
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== bench no-graph"
timeout 600 python bench.py --steps 30 --warmup 5 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.log 2>&1; tail -c 600 gpurun_out/bench_nograph.log
if [ "$1" == "ncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full capture of the K/V projection GEMM"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 8 -c 3 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 4 -c 2 -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_attn.log 2>&1; echo "ncu attn rc=$?"
fi
