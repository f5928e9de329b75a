# round-end verification on one B200: full GPU suite, smoke, the driver's bench command; outputs under gpurun_out/
set -x
(timeout 1200 python -m pytest tests -m gpu -q) > gpurun_out/r02_pytest_gpu.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
PQ3D_BENCH_WATCHDOG=280 timeout 330 python bench.py --steps 100 --warmup 5 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; grep -v "UserWarning\|run_backward" gpurun_out/r02_bench_c3.err | tail -5; cut -c1-250 gpurun_out/r02_bench_c3.json
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_c3_reference.json 2>/dev/null; cut -c1-250 gpurun_out/r02_bench_c3_reference.json
for c in c1 c2 c4; do timeout 200 python bench.py --workload $c --no-cpu-baseline --no-extras --steps 50 --warmup 5 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err; cut -c1-160 gpurun_out/r02_bench_$c.json; done
