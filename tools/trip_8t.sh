set -x
for v in "PQ3D_GRAD_OVERLAP=0"; do
env $v PQ3D_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --train --workload c5 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/tmp_train8.json 2> gpurun_out/tmp_train8.err
python - <<PY
import json
d=json.load(open("gpurun_out/tmp_train8.json"))
print("8 GPUs $v: ms", round(d["ms_per_step"],3), "no-comm", round(d["comm"]["ms_per_step_without_allreduce"],3), "exposed", round(d["comm"]["exposed_comm_ms"],3), "buckets", d["comm"]["buckets"], "value", round(d["value"]))
PY
done
