set -x
PQ3D_BENCH_WATCHDOG=280 timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_8gpu.json 2> gpurun_out/r02_bench_c3_8gpu.err; grep -v "UserWarning\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/r02_bench_c3_8gpu.err | tail -8 | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_c3_8gpu.json"))
print("8gpu value", round(d["value"]), "ms", d["ms_per_step"], "serial", d["serial"])
print("e2e", d["e2e"]["value"], "dec", d["e2e"]["decoder_boundary"]["value"])
print("sustained", d["sustained"]["value"], d["sustained"]["clocks"])
print("train", d["train"])
PY
nvidia-smi topo -m | head -12
