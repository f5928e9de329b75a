set -x
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_nccl_buckets.py > gpurun_out/r02_nccl_buckets.log 2>&1; grep -E "nccl in-backward|Error|assert|Traceback" -A3 gpurun_out/r02_nccl_buckets.log | head -30
for v in "PQ3D_GRAD_OVERLAP=1" "PQ3D_GRAD_OVERLAP=0" "PQ3D_GRAD_OVERLAP=1 NCCL_MAX_CTAS=4" "PQ3D_GRAD_OVERLAP=0 NCCL_MAX_CTAS=8"; do
env $v PQ3D_BENCH_WATCHDOG=200 timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --train --workload c5 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/tmp_train.json 2> gpurun_out/tmp_train.err
python - <<PY
import json
d=json.load(open("gpurun_out/tmp_train.json"))
print("$v: ms", round(d["ms_per_step"],3), "no-comm", round(d["comm"]["ms_per_step_without_allreduce"],3), "exposed", round(d["comm"]["exposed_comm_ms"],3), "buckets", d["comm"]["buckets"], d["comm"]["overlapped_with_backward"])
PY
done
