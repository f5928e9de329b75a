"""timeout 120 torchrun --nproc-per-node 2 tools/check_nccl_grad_mean.py — pq3d_b200.dist.FlatGradAllReduce's flat fp32
path (no encoder given) against an all_gather + mean, eagerly and captured in a CUDA graph.  The in-backward bf16 buckets
are checked by tools/check_nccl_buckets.py.  ALWAYS run under `timeout`; the script also arms its own 90 s alarm."""
import signal

signal.alarm(90)
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pq3d_b200.dist import FlatGradAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device="cpu").manual_seed(100 + rank)
shapes = [(2304, 768), (768,), (12, 5), (2048, 768), (7,)]
params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in shapes]
grads = [torch.randn(s, generator=g).to(dev) for s in shapes]
red = FlatGradAllReduce(params)
for mode in ("eager", "graph"):
    for p, gr in zip(params, grads):
        p.grad = gr.clone()
    if mode == "eager":
        red()
    else:
        red()                                   # warm-up (communicator set-up) outside the capture
        for p, gr in zip(params, grads):
            p.grad.copy_(gr)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            red()
        for p, gr in zip(params, grads):
            p.grad.copy_(gr)
        cg.replay()
    torch.cuda.synchronize()
    for p, gr in zip(params, grads):
        bufs = [torch.empty_like(gr) for _ in range(world)]
        dist.all_gather(bufs, gr)
        ref = torch.stack(bufs).mean(0)
        err = (p.grad - ref).abs().max().item()
        assert err <= 1e-6, (mode, tuple(gr.shape), err)
    if rank == 0:
        print(f"nccl gradient mean ({mode}): OK over {len(shapes)} tensors, world {world}")
dist.destroy_process_group()
