set -x
(timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -s) > gpurun_out/r2g_kernels.log 2>&1; tail -4 gpurun_out/r2g_kernels.log
(timeout 900 python -m pytest tests/test_decoder_gpu.py -m gpu -q -x) > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
PQ3D_GEMM_MULTICAST=0 timeout 200 python tools/chainbench.py 2>&1 | grep linear > gpurun_out/r2g_chain_mc0.txt
PQ3D_GEMM_MULTICAST=1 timeout 200 python tools/chainbench.py 2>&1 | grep -E "linear|out-proj" > gpurun_out/r2g_chain_mc1.txt
paste -d'\n' gpurun_out/r2g_chain_mc0.txt gpurun_out/r2g_chain_mc1.txt
for mc in 0 1; do
  PQ3D_GEMM_MULTICAST=$mc timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/r2g_bench_mc$mc.json 2> gpurun_out/r2g_bench_mc$mc.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench_mc$mc.json"))
print("mc=$mc value", round(d["value"]), "ms", round(d["ms_per_step"],4), "serial", round(d["serial"]["ms_per_step"],4))
PY
done
