set -x
(timeout 900 python -m pytest tests/test_decoder_gpu.py tests/test_kernels_gpu.py -m gpu -q -x) > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
for ctas in 0 100 116 84; do
  if [ $ctas = 0 ]; then export PQ3D_KV_OVERLAP=0; else export PQ3D_KV_OVERLAP=1 PQ3D_KV_OVERLAP_CTAS=$ctas; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r2e_bench_$ctas.json 2> gpurun_out/r2e_bench_$ctas.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2e_bench_$ctas.json"))
print("ctas=$ctas value", round(d["value"]), "ms", round(d["ms_per_step"],4), "serial", d.get("serial"), "e2e", d.get("e2e",{}).get("value"))
PY
done
unset PQ3D_KV_OVERLAP PQ3D_KV_OVERLAP_CTAS
(timeout 1500 python -m pytest tests/test_parity_matched_gpu.py -m gpu -q -s) > gpurun_out/r2e_matched.log 2>&1; grep -E "^c[1-4]|passed|failed|Error|assert" gpurun_out/r2e_matched.log | cut -c1-400 | tail -30
