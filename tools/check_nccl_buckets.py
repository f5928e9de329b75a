"""timeout 150 torchrun --nproc-per-node 2 tools/check_nccl_buckets.py — the in-backward bf16 gradient buckets
(pq3d_b200.dist.FlatGradAllReduce with encoder=...) on the CUDA kernels over NCCL: every rank must end with the mean of the
per-rank gradients (bf16 wire tolerance).  ALWAYS run under `timeout`; the script also arms its own alarm."""
import signal

signal.alarm(120)
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pq3d_b200 import synth  # noqa: E402
from pq3d_b200.dist import FlatGradAllReduce  # noqa: E402
from pq3d_b200.query_encoder import QueryMaskEncoder  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = synth.Workload("g", 2, 100, 512, ["mv", "pc", "voxel", "prompt"], "mixed", T=16, num_layers=2, seed=7)
sd = synth.decoder_state_dict(w, seed=3)


def grads(reduce):
    enc = QueryMaskEncoder(None, **w.decoder_kwargs())
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(dev).train()
    enc.train_dropout = 0.0
    red = FlatGradAllReduce(list(enc.parameters()), encoder=enc) if reduce else None
    inp, pw, _ = synth.make_decoder_inputs(w, rank=rank, device=dev)
    out = enc(synth.clone_input_dict(inp), pw)[0]
    (out ** 2).mean().backward()
    if red is not None:
        left = len([p for p in red.params if id(p) not in red._reduced])
        assert left == 0 or not red.overlapped, left
        red()
    torch.cuda.synchronize()
    return {n: p.grad.clone() for n, p in enc.named_parameters()}, red


local_g, _ = grads(False)
avg, red = grads(True)
worst = 0.0
for n, g in local_g.items():
    parts = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(parts, g)
    want = sum(parts) / world
    worst = max(worst, float((avg[n] - want).abs().max() / want.abs().max().clamp_min(1e-12)))
if rank == 0:
    print(f"nccl in-backward buckets: {red.n_buckets} buckets, {red.bytes_per_step / 1e6:.1f} MB {red.wire_dtype} per step, "
          f"worst rel err vs all_gather mean {worst:.2e}")
assert worst <= 2e-2, worst
dist.destroy_process_group()
