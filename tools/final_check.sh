set -x
(timeout 700 python -m pytest tests -m gpu -x -q) > gpurun_out/r1g_pytest.log 2>&1; tail -4 gpurun_out/r1g_pytest.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/r1g_bench_c3.json 2> gpurun_out/r1g_bench_c3.err; cut -c1-220 gpurun_out/r1g_bench_c3.json; tail -2 gpurun_out/r1g_bench_c3.err
timeout 100 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r1g_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r1g_bench_ref.json
