"""CPU, gloo, world_size 2: the host logic of the N>1 path — scene sharding, ordered gather,
max-over-ranks timing, flat gradient mean — and that sharded oracle outputs equal the unsharded
ones (scenes are independent, so batch-axis sharding needs no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pq3d_b200 import dist as pd
from pq3d_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def test_balanced_shards_cover_and_balance():
    toks = [4096, 128, 3000, 2000, 900, 3500, 256, 1024]
    for world in (1, 2, 4, 8):
        shards = pd.balanced_scene_shards(toks, world)
        assert sorted(i for s in shards for i in s) == list(range(len(toks)))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    s2 = pd.balanced_scene_shards(toks, 2)
    loads = [sum(toks[i] for i in s) for s in s2]
    assert abs(loads[0] - loads[1]) <= 0.15 * sum(toks)
    assert pd.balanced_scene_shards([5, 5, 5], 2) == [[0, 2], [1]]


def _gather_job(rank, world):
    toks = [300, 100, 200, 50, 250]
    shards = pd.balanced_scene_shards(toks, world)
    full = torch.arange(5 * 3, dtype=torch.float32).view(5, 3)
    local = full[shards[rank]] * 1.0
    out = pd.gather_in_order(local, shards)
    t = pd.max_over_ranks(1.0 + rank)
    return bool(torch.equal(out, full)), t


def test_gather_in_order_and_max_timing():
    r = _spawn(_gather_job)
    assert all(v[0] for v in r.values())
    assert all(v[1] == 2.0 for v in r.values())


def _grad_job(rank, world):
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    unused = torch.nn.Parameter(torch.zeros(2))
    params = list(lin.parameters()) + [unused]
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    ar = pd.FlatGradAllReduce(params)
    ar()
    return [p.grad.clone() for p in params]


def test_flat_grad_allreduce_means_and_fills_unused():
    r = _spawn(_grad_job)
    g0, g1 = r[0], r[1]
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)
    # weight grad of sum(lin(x)) is sum_batch x: rank0 -> 2*1, rank1 -> 2*2; mean = 3
    assert torch.allclose(g0[0], torch.full((3, 4), 3.0))
    assert torch.equal(g0[2], torch.zeros(2))


def _shard_oracle_job(rank, world):
    from oracle import restatement as O
    w = synth.Workload("t", 4, 12, 40, ["mv", "pc"], "parallel", num_layers=1, ragged=(16, 40))
    sd = synth.decoder_state_dict(w, seed=3)
    inp, pw, d = synth.make_decoder_inputs(w)
    cfg = O.DecoderCfg(**w.decoder_kwargs())
    toks = d["seg_pad_masks"].sum(1).tolist()
    shards = pd.balanced_scene_shards(toks, world)
    ix = torch.tensor(shards[rank])
    sub = {"query": tuple(t[ix] for t in inp["query"])}
    for m in ("mv", "pc"):
        sub[m] = [t[ix] for t in inp[m]]
    with torch.no_grad():
        local = O.query_mask_encoder(sd, cfg, sub, pw[ix])[0]
        full = O.query_mask_encoder(sd, cfg, synth.clone_input_dict(inp), pw)[0]
    out = pd.gather_in_order(local, shards)
    return float((out - full).abs().max())


def test_sharded_equals_unsharded():
    r = _spawn(_shard_oracle_job)
    assert all(v <= 1e-5 for v in r.values()), r


def test_take_scenes():
    w = synth.Workload("t", 3, 4, 8, ["voxel"], "parallel", voxel_multiscale=True, num_layers=2)
    d = synth.make_data_dict(w)
    s = pd.take_scenes(d, [2, 0])
    assert s["seg_center"].shape[0] == 2 and torch.equal(s["seg_center"][0], d["seg_center"][2])
    assert isinstance(s["voxel_seg_fts_multiscale"], list) and s["voxel_seg_fts_multiscale"][0].shape[0] == 2


def _bucket_job(rank, world):
    """In-backward bf16 gradient buckets (dist.FlatGradAllReduce with encoder=...) on the emulated kernels: every rank
    ends with the mean of the per-rank gradients, and nothing is left for the flat end-of-backward path."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _cpu_ops
    from pq3d_b200.query_encoder import QueryMaskEncoder
    w = synth.Workload("g", 2, 12, 40, ["mv", "pc", "prompt"], "mixed", T=5, num_layers=2, seed=7)
    sd = synth.decoder_state_dict(w, seed=3)

    def grads(reduce):
        enc = QueryMaskEncoder(None, **w.decoder_kwargs())
        enc.load_state_dict(sd, strict=True)
        enc.use_cuda_graph, enc.train_streams, enc.train_dropout = False, False, 0.0
        enc.train()
        red = pd.FlatGradAllReduce(list(enc.parameters()), encoder=enc) if reduce else None
        inp, pw, _ = synth.make_decoder_inputs(w, rank=rank)              # this rank's scenes
        with _cpu_ops.cpu_backend():
            out = enc(synth.clone_input_dict(inp), pw)[0]
            (out ** 2).mean().backward()
        left = None
        if red is not None:
            left = len([p for p in red.params if id(p) not in red._reduced])
            red()
        return {n: p.grad.clone() for n, p in enc.named_parameters()}, red, left
    local, _, _ = grads(False)
    avg, red, left = grads(True)
    worst = 0.0
    for n, g in local.items():
        parts = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        want = sum(parts) / world
        scale = want.abs().max().clamp_min(1e-12)
        worst = max(worst, float((avg[n] - want).abs().max() / scale))
    return worst, left, red.overlapped, red.n_buckets, red.wire_dtype


def test_in_backward_bf16_buckets_average_gradients():
    r = _spawn(_bucket_job)
    for worst, left, overlapped, n_buckets, wire in r.values():
        assert overlapped and left == 0 and n_buckets == 2 + 3 + 1 and wire == "bfloat16"
        assert worst <= 1.5e-2, worst                 # bf16 on the wire: 2^-8 per addend
