set -x
(timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_parity_matched_gpu.py -m gpu -q -s) > gpurun_out/r2h_pytest.log 2>&1; grep -E "^c[1-4]|passed|failed|FAILED" gpurun_out/r2h_pytest.log | cut -c1-500
for sb in 0 1; do
  PQ3D_SIDE_BRANCHES=$sb timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 200 > gpurun_out/r2h_bench_sb$sb.json 2> gpurun_out/r2h_bench_sb$sb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench_sb$sb.json"))
print("side_branches=$sb value", round(d["value"]), "ms", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "attn us", d["roofline_attention"]["us_per_launch"], d["roofline_attention"]["us_per_launch_l2_warm"])
PY
done
PQ3D_PDL=0 timeout 150 python tools/timeline.py c3 > gpurun_out/timeline2_pdl0.txt 2>&1
timeout 600 python bench.py --steps 100 > gpurun_out/r2h_bench_full.json 2> gpurun_out/r2h_bench_full.err; tail -3 gpurun_out/r2h_bench_full.err; cut -c1-400 gpurun_out/r2h_bench_full.json
