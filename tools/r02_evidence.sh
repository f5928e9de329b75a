# round-2 evidence set (one B200): ncu --set full of the two hot kernels + the chain GEMM, serialised launch list of one
# inference step, plain launch list of the training step.  Numbers under ncu are never bench values.
set -x
ncu --set full --clock-control none --import-source on -k regex:attention_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_attn python tools/ncu_targets.py attn > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_bf16_kernel -s 3 -c 1 -o gpurun_out/r02_gemm python tools/ncu_targets.py gemm > /dev/null 2>&1
PQ3D_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r02_ncu_launches_c3.csv python bench.py --steps 3 --warmup 5 --no-graph --streams 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la gpurun_out/r02_* | head
