set -x
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r1f_pytest.log 2>&1; tail -3 gpurun_out/r1f_pytest.log
timeout 300 python bench.py > gpurun_out/r1f_bench_c3.json 2> gpurun_out/r1f_bench_c3.err; cut -c1-200 gpurun_out/r1f_bench_c3.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1f_bench_c3_reference.json 2>> gpurun_out/r1f_bench_c3.err; cut -c1-200 gpurun_out/r1f_bench_c3_reference.json
for c in c1 c2 c4; do timeout 200 python bench.py --workload $c --no-cpu-baseline --steps 50 > gpurun_out/r1f_bench_$c.json 2> gpurun_out/r1f_bench_$c.err; cut -c1-160 gpurun_out/r1f_bench_$c.json; done
timeout 300 python bench.py --train --workload c5 > gpurun_out/r1f_bench_c5.json 2> gpurun_out/r1f_bench_c5.err; cut -c1-200 gpurun_out/r1f_bench_c5.json
python tools/train_profile.py c5 > gpurun_out/r1f_train_profile.txt 2>&1
python tools/profile_step.py c3 > gpurun_out/r1f_step_profile.txt 2>&1
PQ3D_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r1f_ncu_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-graph --streams 1 --no-cpu-baseline > /dev/null 2>&1
PQ3D_PDL=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 600 --csv --log-file gpurun_out/r1f_ncu_launches_c5.csv python bench.py --train --workload c5 --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:linear_bf16_kernel -s 5 -c 1 -o gpurun_out/r1f_gemm python tools/ncu_targets.py gemm > /dev/null 2>&1
ls -la gpurun_out | tail -20
