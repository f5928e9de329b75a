"""Whole-step CUDA graph for training: forward + loss + backward + gradient all-reduce + optimizer step replayed as
ONE graph launch.

A training step of the decoder issues ~600 kernels of a few microseconds each (the query side is 400 rows); launched
one by one from Python the step is host-bound (measured on B200, BASELINE config 5 shard: 15.9 ms/step eager against
9.6 ms of GPU time).  The reference's loop (trainer/query3d_trainer.py:18-28: forward, loss, backward, optimizer.step)
maps onto `GraphedTrainStep.__call__`; the optimizer must be constructed with `capturable=True`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def _map_tensors(obj, f, memo):
    if isinstance(obj, torch.Tensor):
        if id(obj) not in memo:
            memo[id(obj)] = f(obj)
        return memo[id(obj)]
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(x, f, memo) for x in obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, f, memo) for k, v in obj.items()}
    return obj


def _copy_into(dst, src, seen):
    if isinstance(dst, torch.Tensor):
        if id(dst) not in seen:
            seen.add(id(dst))
            dst.copy_(src, non_blocking=True)
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_into(d, s, seen)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k], seen)


class GraphedTrainStep:
    """step(input_dict, pairwise_locs, *loss_args) -> loss (a static device tensor, valid until the next call).

    encoder      pq3d_b200.QueryMaskEncoder (training path: `train_dropout = 0.0`)
    optimizer    torch optimizer built with capturable=True
    loss_fn      loss_fn(decoded_queries, *loss_args) -> scalar tensor
    reducer      optional callable run between backward and optimizer.step (pq3d_b200.dist.FlatGradAllReduce)

    pre_step     optional callable run after the reducer and before optimizer.step, inside the captured step — the hook
                 for the reference's `clip_grad_norm_` (trainer/query3d_trainer.py:24); it must be capturable (torch's
                 foreach clip_grad_norm_ with error_if_nonfinite=False is)

    The first `warmup` calls run eagerly on a side stream (allocator / kernel-attribute warm-up), the next call
    captures, later calls copy the new batch into the captured input buffers and replay.  Shapes must not change.

    Learning rate: a CUDA graph freezes Python floats.  Build the optimizer with `lr=torch.tensor(lr, device=...)`
    (torch's capturable optimizers read a tensor lr on the device) and let the scheduler write into that tensor
    (`set_lr`), otherwise the reference's warm-up / cosine LambdaLR has no effect on replayed steps.  A float lr is
    converted to a device tensor here for exactly that reason.
    """

    def __init__(self, encoder, optimizer, loss_fn: Callable, reducer: Optional[Callable] = None, warmup: int = 3,
                 pre_step: Optional[Callable] = None):
        self.enc, self.opt, self.loss_fn, self.reducer = encoder, optimizer, loss_fn, reducer
        self.pre_step = pre_step
        dev = next(encoder.parameters()).device
        for grp in optimizer.param_groups:            # float lr -> device tensor, so schedulers act on replayed steps
            if not isinstance(grp.get("lr"), torch.Tensor) and dev.type == "cuda":
                grp["lr"] = torch.tensor(float(grp["lr"]), dtype=torch.float32, device=dev)
        self.warmup, self.calls = max(1, warmup), 0   # >= 1: buffers the capture reuses are created eagerly
        self.graph = None
        self.static = None
        self.loss = None
        self.launches = 0
        self._side = None

    def _eager(self, inp, pw, loss_args):
        self.opt.zero_grad(set_to_none=True)
        out = self.enc(self._clone_struct(inp), pw)[0]
        loss = self.loss_fn(out, *loss_args)
        loss.backward()
        if self.reducer is not None:
            self.reducer()
        if self.pre_step is not None:
            self.pre_step()
        self.opt.step()
        return loss

    def set_lr(self, lr: float, group: Optional[int] = None):
        """Write a new learning rate into the optimizer's device-side lr tensors (takes effect on the next replay)."""
        for gi, grp in enumerate(self.opt.param_groups):
            if group is None or gi == group:
                if isinstance(grp["lr"], torch.Tensor):
                    grp["lr"].fill_(float(lr))
                else:
                    grp["lr"] = float(lr)

    @staticmethod
    def _clone_struct(inp):
        return {k: (tuple(v) if isinstance(v, tuple) else list(v)) for k, v in inp.items()}

    def __call__(self, input_dict, pairwise_locs, *loss_args):
        from . import ops
        if self.graph is not None:
            _copy_into(self.static, (input_dict, pairwise_locs, loss_args), set())
            self.graph.replay()
            ops._count(self.launches)
            self.enc.mark_weights_changed()  # the packed operand copies trail the parameters by one step
            return self.loss
        if self.calls < self.warmup:
            self.calls += 1
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                loss = self._eager(input_dict, pairwise_locs, loss_args)
            torch.cuda.current_stream().wait_stream(self._side)
            return loss
        memo = {}
        self.static = _map_tensors((input_dict, pairwise_locs, loss_args), lambda t: t.clone(), memo)
        inp, pw, largs = self.static
        torch.cuda.synchronize()
        self.opt.zero_grad(set_to_none=True)
        before = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            out = self.enc(self._clone_struct(inp), pw)[0]
            self.loss = self.loss_fn(out, *largs)
            self.loss.backward()
            if self.reducer is not None:
                self.reducer()
            if self.pre_step is not None:
                self.pre_step()
            self.opt.step()
        self.launches = ops.LAUNCHES - before
        ops.LAUNCHES = before
        self.graph.replay()
        ops._count(self.launches)
        self.enc.mark_weights_changed()
        return self.loss
