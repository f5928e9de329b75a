"""Tensor restatement of the kernels' counter-based dropout RNG (csrc/ptx.cuh: hash32 / drop_key / drop_keep), so
tests can rebuild the exact keep masks a training step used and hand them to the oracle.  Pure torch integer ops; runs
on any device.  Site numbering (one independent stream per dropout call site and layer) lives here too."""
from __future__ import annotations

import torch

M32 = 0xFFFFFFFF
SITES_PER_LAYER = 16
SITE_CA_SUBLAYER = 0      # + group index (0..3): dropout on a cross-attention group's out-projection, idx = (g*R+row)*D+col
SITE_SA_SUBLAYER = 4
SITE_FFN_SUBLAYER = 5
SITE_FFN_HIDDEN = 6       # idx = row*F + col
SITE_SA_PROBS = 7         # plain self-attention probabilities, idx = ((b*H+h)*N + n)*ceil128(S) + key
SITE_CA_PROBS = 8         # + memory index in `memories` order
SITE_MASK_HEAD = 1 << 16  # + call index: the cls head's Dropout on LN(relu(linear0 q)), idx = row*hidden + col


def site(layer: int, kind: int) -> int:
    return layer * SITES_PER_LAYER + kind


def hash32(x: torch.Tensor) -> torch.Tensor:
    x = x & M32
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & M32
    x = x ^ (x >> 16)
    return x


def threshold(p: float) -> int:
    t = float(torch.tensor(p, dtype=torch.float32).item()) * 4294967296.0     # p arrives at the kernels as a C float
    return 4294967295 if t >= 4294967295.0 else int(t)


def keep_mask(seed: int, site_id: int, idx: torch.Tensor, p: float) -> torch.Tensor:
    """bool tensor shaped like idx (int64 element indices): True = kept."""
    key = hash32(torch.tensor((seed + site_id * 0x9E3779B9) & M32, dtype=torch.int64, device=idx.device))
    return hash32((idx.to(torch.int64) & M32) ^ key) >= threshold(p)
