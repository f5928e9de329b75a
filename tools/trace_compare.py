"""Debug aid: run the decoder once on the CUDA kernels and once on tests/_cpu_ops.py's emulation, recording every tensor
argument of every `ops.*` call after the call, and print where the two traces start to differ (max-norm relative)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _cpu_ops
from pq3d_b200 import ops, synth
from pq3d_b200.query_encoder import QueryMaskEncoder

NAMES = ["linear", "attention", "spatial_bias", "ingest_memory", "add_layernorm", "pack_mask", "cast_bf16", "gate_mix"]


def tensors_of(args, kwargs):
    out = []
    def walk(tag, v):
        if isinstance(v, torch.Tensor):
            out.append((tag, v.detach().float().cpu().clone()))
        elif isinstance(v, ops.AttnMemory):
            pass
        elif isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                walk(f"{tag}[{i}]", x)
    for i, a in enumerate(args):
        walk(f"a{i}", a)
    for k, a in kwargs.items():
        walk(k, a)
    return out


def record(trace):
    saved = {}
    for n in NAMES:
        f = getattr(ops, n)
        saved[n] = f
        def wrap(*a, _f=f, _n=n, **kw):
            r = _f(*a, **kw)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            trace.append((_n, tensors_of(a, kw), {k: v for k, v in kw.items() if isinstance(v, (int, float, bool))}))
            return r
        setattr(ops, n, wrap)
    return saved


def restore(saved):
    for n, f in saved.items():
        setattr(ops, n, f)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c1"
    w = synth.workload(name)
    if len(sys.argv) > 2:
        w.num_layers = int(sys.argv[2])
    sd = synth.decoder_state_dict(w, seed=0, sharp=2.0)
    gpu = QueryMaskEncoder(None, **w.decoder_kwargs()).eval(); gpu.load_state_dict(sd); gpu = gpu.cuda(); gpu.use_cuda_graph = False
    cpu = QueryMaskEncoder(None, **w.decoder_kwargs()).eval(); cpu.load_state_dict(sd); cpu.use_cuda_graph = False
    inp, pw, d = synth.make_decoder_inputs(w, device="cuda")
    tg, tc = [], []
    s = record(tg)
    with torch.no_grad():
        gpu(synth.clone_input_dict(inp), pw)
    restore(s)
    to_cpu = lambda x: x.cpu() if isinstance(x, torch.Tensor) else (type(x)(to_cpu(v) for v in x) if isinstance(x, (list, tuple)) else x)
    inp_c = {k: to_cpu(v) for k, v in inp.items()}
    with _cpu_ops.cpu_backend():
        s = record(tc)
        with torch.no_grad():
            cpu(synth.clone_input_dict(inp_c), pw.cpu() if pw is not None else None)
        restore(s)
    print(len(tg), len(tc))
    for i, ((ng, ag, kg), (nc, ac, kc)) in enumerate(zip(tg, tc)):
        assert ng == nc, (ng, nc)
        parts = []
        for (tag, a), (_, b) in zip(ag, ac):
            if a.shape != b.shape:
                parts.append(f"{tag}:shape"); continue
            fin = torch.isfinite(b) & torch.isfinite(a)
            den = b[fin].abs().max().clamp_min(1e-30) if fin.any() else torch.tensor(1.0)
            e = ((a - b).abs()[fin].max() / den).item() if fin.any() else 0.0
            nd = int(((a != b) & fin).sum())
            parts.append(f"{tag}:{e:.1e}({nd})")
        print(i, ng, {k: v for k, v in kg.items() if k in ("M", "N", "K", "G", "relu")}, " ".join(parts))


main()
