"""Multi-GPU plumbing: one process per GPU, scenes sharded over the batch axis.

The reference's only parallelism is DDP through HF accelerate (trainer/build.py:66-75,123-129): the
batch is split across ranks, weights are replicated, and the sole collective that follows the hot path
is the gradient all-reduce in backward.  No op of the decoder mixes information across scenes
(SURVEY.md §8e), so inference needs no data-path collective at all; the helpers here are the host
logic around that: which scenes a rank owns (balanced by segment-token count for ragged batches,
since cost is linear in S), putting gathered results back in order, max-over-ranks timing, and the
flat-buffer gradient mean for the training step.  Works on NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def balanced_scene_shards(tokens_per_scene: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time greedy: assign scenes (heaviest first) to the lightest rank, then keep
    each rank's scene ids sorted.  Every rank gets len/world scenes +-1 when counts allow."""
    n = len(tokens_per_scene)
    cap = -(-n // world)
    order = sorted(range(n), key=lambda i: (-tokens_per_scene[i], i))
    load = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        cands = [r for r in range(world) if len(shards[r]) < cap]
        r = min(cands, key=lambda r: (load[r], r))
        shards[r].append(i)
        load[r] += tokens_per_scene[i]
    return [sorted(s) for s in shards]


def take_scenes(data_dict: Dict, idx: Sequence[int]) -> Dict:
    """Batch-axis slice of a collated data_dict (tensors, or lists of tensors for multi-scale voxels)."""
    ix = torch.as_tensor(list(idx), dtype=torch.long)

    def f(v):
        if isinstance(v, torch.Tensor):
            return v.index_select(0, ix.to(v.device))
        if isinstance(v, (list, tuple)):
            return type(v)(f(x) for x in v)
        return v
    return {k: f(v) for k, v in data_dict.items()}


def gather_in_order(local: torch.Tensor, shards: List[List[int]], group=None) -> torch.Tensor:
    """All-gather per-rank results (first dim = that rank's scenes, possibly uneven) and restore the
    original scene order.  Used by evaluation / tests, never inside the timed decoder path."""
    world = dist.get_world_size(group)
    cap = max(len(s) for s in shards)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    n = sum(len(s) for s in shards)
    out = torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, s in enumerate(shards):
        for j, scene in enumerate(s):
            out[scene] = bufs[r][j]
    return out


def max_over_ranks(value: float, device="cpu", group=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.item()


class FlatGradAllReduce:
    """One flat buffer for all gradients -> a single all-reduce (mean) per step: on an NVSwitch box the collective is
    latency-, not link-bound, so one ~240 MB fp32 call beats DDP's 25 MB buckets.  Per step: ONE multi-tensor copy of
    the gradients into the flat buffer, the all-reduce, one division, and then `p.grad` is re-pointed at its view of
    the flat buffer (no copy back; the optimizer reads the averaged gradients in place).  The first version issued
    three small kernels per parameter (~500 launches, +1.5 ms on a 3.2 ms step at 2 GPUs).  Capturable in the training
    step's CUDA graph.  Parameters without a gradient this step contribute zeros (the reference runs DDP with
    find_unused_parameters=True, trainer/build.py:66).

    PQ3D_COALESCED_ALLREDUCE=1 selects an in-place variant (every gradient all-reduced inside one ncclGroup through
    torch's coalescing manager, no staging copy at all); it is experimental: not verified under CUDA-graph capture."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        import os
        self.coalesced = os.environ.get("PQ3D_COALESCED_ALLREDUCE", "0") == "1"

    def __call__(self):
        world = dist.get_world_size(self.group)
        if self.coalesced and dist.get_backend(self.group) == "nccl":
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            with dist._coalescing_manager(group=self.group, device=self.params[0].device, async_ops=False):
                for p in self.params:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.group)
            return
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            p.grad = v
