set -x
(timeout 1200 python -m pytest tests -m gpu -q -x -s -k "matched") > gpurun_out/r2a_matched.log 2>&1; tail -15 gpurun_out/r2a_matched.log
(timeout 900 python -m pytest tests -m gpu -q -k "not matched") > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; cut -c1-300 gpurun_out/r2a_bench_c3.json
