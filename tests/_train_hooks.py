"""Shared by the GPU and the CPU (emulated-kernel) training tests: the oracle hook that replays the kernels' dropout
masks."""
import torch


class KernelRngTrain:
    """cfg.train hook for the oracle that rebuilds the kernels' keep masks from (seed, site, element index) with the
    tensor restatement of the counter RNG (pq3d_b200/rng.py) — element indexing as documented in include/pq3d_b200.h."""

    def __init__(self, enc, seed, p, B, N, H):
        from pq3d_b200 import rng
        self.rng, self.enc, self.seed, self.p, self.B, self.N, self.H = rng, enc, seed, p, B, N, H
        self.program = [g for g in enc._program() if len(g) > 0]
        self.keeps = list(enc.last_memory_keep)

    def _apply(self, x, keep):
        return x * keep.to(x.dtype) / (1.0 - self.p)

    def sublayer(self, layer, kind, x):
        rng, (B, N, D) = self.rng, x.shape
        R = B * N
        e = torch.arange(R * D, device=x.device).view(B, N, D)
        if isinstance(kind, tuple):
            gi = next(i for i, g in enumerate(self.program) if kind[1] in g)
            e = e + self.program[gi].index(kind[1]) * R * D
            site = rng.site(layer, rng.SITE_CA_SUBLAYER + gi)
        else:
            site = rng.site(layer, rng.SITE_SA_SUBLAYER if kind == "sa" else rng.SITE_FFN_SUBLAYER)
        return self._apply(x, rng.keep_mask(self.seed, site, e, self.p))

    def probs(self, layer, kind, P):
        rng, H = self.rng, self.H
        BH, L, S2 = P.shape
        S = S2 - 1 if isinstance(kind, tuple) else S2            # cross-attention carries the zero-attn column
        s_pad = (S + 127) // 128 * 128
        rows = torch.arange(BH * L, device=P.device).view(BH, L, 1)          # (b*H + h)*N + n
        e = rows * s_pad + torch.arange(S, device=P.device).view(1, 1, S)
        site = (rng.site(layer, rng.SITE_CA_PROBS + self.enc.memories.index(kind[1])) if isinstance(kind, tuple)
                else rng.site(layer, rng.SITE_SA_PROBS))
        keep = rng.keep_mask(self.seed, site, e, self.p)
        if S2 != S:
            keep = torch.cat([keep, torch.ones_like(keep[..., :1])], -1)
        return self._apply(P, keep)

    def hidden(self, layer, h):
        e = torch.arange(h.numel(), device=h.device).view(h.shape)
        return self._apply(h, self.rng.keep_mask(self.seed, self.rng.site(layer, self.rng.SITE_FFN_HIDDEN), e, self.p))

    def memory_keep(self, layer, memories, B):
        return self.keeps.pop(0)
