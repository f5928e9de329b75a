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
    """Gradient mean over the ranks for the training step — the one collective of the path (the reference: DDP through
    accelerate, trainer/build.py:66-75, whose buckets overlap backward by construction).

    Two cooperating parts:

    * IN-BACKWARD BUCKETS (when constructed with `encoder=` and the world size is <= 2 — measured, see __init__ — or
      PQ3D_GRAD_OVERLAP=1; the decoder is the one whose hand-composed backward,
      pq3d_b200/train_engine.py, calls back): as soon as the backward of decoder layer i has produced that layer's
      query-side gradients (self-attention, FFN, out-projections, LayerNorms — final once `group_bwd` of the layer
      returns), they are packed as bf16 into one bucket and all-reduced on a COMMUNICATION STREAM while the backward of
      layer i-1 runs; the in-projection gradients (their K / V rows are only final in the memory-side tail) go out
      per memory, as soon as that memory's weight-gradient GEMMs are queued, so the next memory's work hides them; what
      is left forms a last, small bucket.  The backward joins the communication stream before it hands the (already averaged) gradients
      to autograd.  bf16 on the wire halves the bytes (120 MB instead of 240 MB per step for the 60 M-parameter
      decoder; the same compression torch's `bf16_compress_hook` applies to DDP buckets); the sum over ranks is
      upcast to fp32 and divided by the world size on arrival.
    * `__call__()` (between backward and optimizer.step): all-reduces whatever the buckets did not cover — parameters
      outside the decoder, or everything when no encoder was given / `num_blocks > 1` / shared layers — through one
      flat fp32 buffer, and re-points `p.grad` at its view (no copy back).

    Parameters without a gradient this step contribute zeros (the reference runs DDP with
    find_unused_parameters=True, trainer/build.py:66).  Capturable in the training step's CUDA graph
    (pq3d_b200/training.py).  `enabled = False` skips every collective (single-rank timing of the same step)."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group=None, encoder=None, wire_dtype=torch.bfloat16):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.enabled = True
        self.wire = wire_dtype
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.dev = dev
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()
        self.encoder = encoder
        self.overlapped = False
        self._name_of = {}
        self._reduced = set()          # ids of parameters whose gradient was averaged inside this step's backward
        self._buckets = {}             # bucket key -> (wire buffer, [(name, offset, numel, shape)])
        self._live = []                # (bucket key, names) launched in the current backward
        self.comm_stream = None
        if encoder is not None:
            self._name_of = {n_: p for n_, p in encoder.named_parameters()}
            ours = {id(p) for p in self.params}
            shared = len({id(p) for p in encoder.parameters()}) != len(list(encoder.named_parameters(remove_duplicate=False)))
            import os
            # MEASURED (config-5 shard, whole step one CUDA graph): 2 GPUs — buckets 3.78 ms/step vs one flat bf16 call
            # 3.89; 8 GPUs — buckets 4.67 vs flat 4.14 (3.18 without any collective).  The persistent GEMMs of the
            # backward need every SM, so a collective running next to them delays them by about its own duration and
            # nine small collectives pay nine latencies at 8 ranks: overlap only pays for two ranks.
            env = os.environ.get("PQ3D_GRAD_OVERLAP")
            world = dist.get_world_size(group) if dist.is_initialized() else 1
            want = (world <= 2) if env is None else (env != "0")
            if want and all(id(p) in ours for p in self._name_of.values()) and getattr(encoder, "num_blocks", 1) == 1 and not shared:
                encoder.grad_sink = self
                self.overlapped = True
                if dev.type == "cuda":
                    self.comm_stream = torch.cuda.Stream(device=dev)
        # the flat path (everything the buckets do not cover; all of it when nothing overlaps) also travels in `wire`
        self.flat_wire = torch.empty(n, dtype=self.wire, device=dev) if (self.wire != torch.float32 and n) else None
        esz = torch.empty(0, dtype=self.wire).element_size()
        self.bytes_per_step = n * esz
        self.wire_dtype = str(self.wire).replace("torch.", "")
        self.n_buckets = (getattr(encoder, "num_layers", 0) + len(getattr(encoder, "memories", [])) + 1) if self.overlapped else 1

    # ---- called by train_engine._Bwd -------------------------------------------------------------------
    def begin_backward(self):
        self._reduced.clear()
        self._live.clear()

    def reduce_async(self, key, grads: Dict[str, torch.Tensor], producers=()):
        """Pack `grads` (name -> final fp32 gradient) into bucket `key` as bf16 and all-reduce it on the communication
        stream, after everything queued so far on the `producers` streams (CUDA) has run."""
        if not self.enabled or not grads:
            return
        names = sorted(grads)
        ent = self._buckets.get(key)
        if ent is None or [e[0] for e in ent[1]] != names:
            layout, off = [], 0
            for n_ in names:
                layout.append((n_, off, grads[n_].numel(), tuple(grads[n_].shape)))
                off += grads[n_].numel()
            ent = self._buckets[key] = (torch.empty(off, dtype=self.wire, device=self.dev), layout)
        buf, layout = ent
        cur = torch.cuda.current_stream(self.dev) if self.comm_stream is not None else None
        if self.comm_stream is not None:
            for st in tuple(producers) + (cur,):
                if st is not None:
                    self.comm_stream.wait_stream(st)
        ctx = torch.cuda.stream(self.comm_stream) if self.comm_stream is not None else _null()
        with ctx:
            torch._foreach_copy_([buf[o:o + k].view(shp) for _, o, k, shp in layout], [grads[n_] for n_ in names])
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        self._live.append(key)

    def finish_backward(self, grads: Dict[str, torch.Tensor]):
        """Join the communication stream and replace the entries of `grads` that went out in buckets by their mean over
        the ranks (fp32).  Called once, at the end of the backward, on the backward's main stream."""
        if not self._live:
            return
        if self.comm_stream is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)
        inv = 1.0 / dist.get_world_size(self.group)
        for key in self._live:
            buf, layout = self._buckets[key]
            avg = buf.float().mul_(inv)
            for n_, o, k, shp in layout:
                grads[n_] = avg[o:o + k].view(shp)
                p = self._name_of.get(n_)
                if p is not None:
                    self._reduced.add(id(p))
        self._live.clear()

    # ---- between backward and optimizer.step -------------------------------------------------------------
    def __call__(self):
        if not self.enabled:
            self._reduced.clear()
            return
        world = dist.get_world_size(self.group)
        rest = [(p, v) for p, v in zip(self.params, self.views) if id(p) not in self._reduced]
        self._reduced.clear()
        if not rest:
            return
        have = [(p, v) for p, v in rest if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for p, v in rest:
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        if len(rest) == len(self.params):
            if self.flat_wire is not None:               # one pass down to the wire type, one collective, one pass back
                self.flat_wire.copy_(self.flat)
                dist.all_reduce(self.flat_wire, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.copy_(self.flat_wire)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
        else:                                    # a subset: reduce its views one by one (rare: parameters outside the decoder)
            for _, v in rest:
                dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.group)
                v.div_(world)
        for p, v in rest:
            p.grad = v


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
