#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_after_bench.csv
export PQ3D_PDL=0   # serialised profiling: no overlap between kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm python tools/ncu_targets.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 3 -c 1 -f -o gpurun_out/prof_attn python tools/ncu_targets.py attn > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
unset PQ3D_PDL
timeout 600 python tools/profile_step.py c3 > gpurun_out/profile_step.log 2>&1; tail -12 gpurun_out/profile_step.log
