"""Training-mode forward + backward of the small dense blocks around the decoder, composed from the sm_100a kernels
(no autograd through PyTorch ops, no fallback):

  linear_ln_train   nn.Sequential(Linear, LayerNorm) — ObjectEncoder.input_feat_proj (modules/vision/object_encoder.py:33-38),
                    CoordinateEncoder.feat_proj and the dim_loc > 3 coordinate / box encoders
                    (model/query3d_unified.py:15-27, 62-69)
  MlpHeadTrain      get_mlp_head: Linear - ReLU - LayerNorm(1e-12) - Dropout - Linear (modules/utils.py:18-25) — the mask
                    head's class branch and GroundHead (modules/heads/grounding_head.py:42-55)

GEMM-shaped work is pq3d_linear_bf16 (dgrad with a transposed bf16 weight, wgrad on K-major transposes); LayerNorm
backward, bias column sums and ReLU / dropout gates are the backward.cu kernels.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops

bf16 = torch.bfloat16
f32 = torch.float32


def pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _tcast(x, rows, cols, want_c=False, gate=None):
    xt = torch.empty(cols, pad64(rows), dtype=bf16, device=x.device)
    xc = torch.empty(rows, cols, dtype=bf16, device=x.device) if want_c else None
    ops.transpose_cast(x, xt, xc, gate=gate)
    return xt, xc


def _wgrad(dyT, xT, n_out, n_in):
    dW = torch.empty(n_out, n_in, dtype=f32, device=dyT.device)
    ops.linear(dyT, xT, dW, M=n_out, N=n_in, K=dyT.shape[1])
    return dW


def _colsum(x, gate=None):
    out = torch.empty(x.shape[1], dtype=f32, device=x.device)
    ops.colsum(x, out, gate=gate)
    return out


# ------------------------------------------------------------------------------------------------------------
# Linear + LayerNorm
# ------------------------------------------------------------------------------------------------------------
class _LinearLNFn(torch.autograd.Function):
    """forward(x or None, x16 or None, weight, bias, ln_weight, ln_bias, eps) -> LayerNorm(x W^T + b), fp32."""

    @staticmethod
    def forward(ctx, x, x16, weight, bias, ln_w, ln_b, eps):
        D, din = weight.shape
        kp = pad64(din)
        dev = weight.device
        w16 = torch.zeros(D, kp, dtype=bf16, device=dev)
        w16[:, :din] = weight.detach()
        if x16 is None:
            lead = x.shape[:-1]
            R = x.numel() // din
            x2 = torch.zeros(R, kp, dtype=f32, device=dev)
            x2[:, :din] = x.detach().reshape(R, din)
            x16 = torch.empty(R, kp, dtype=bf16, device=dev)
            ops.cast_bf16(x2, x16)
        else:
            R, lead = x16.shape[0], (x16.shape[0],)
        y = torch.empty(R, D, dtype=f32, device=dev)
        ops.linear(x16, w16, y, M=R, N=D, K=kp, bias=bias.detach().float().contiguous())
        out = torch.empty(R, D, dtype=f32, device=dev)
        g = ln_w.detach().float().contiguous()[None]
        ops.add_layernorm(y, None, g, ln_b.detach().float().contiguous()[None], eps, R, D, out_f32=out)
        ctx.save_for_backward(x16, y, w16, g)
        ctx.eps, ctx.din, ctx.lead, ctx.x_dtype = eps, din, lead, (None if x is None else x.dtype)
        return out.view(*lead, D)

    @staticmethod
    def backward(ctx, d_out):
        x16, y, w16, g = ctx.saved_tensors
        R, D = y.shape
        kp, din, dev = x16.shape[1], ctx.din, y.device
        d_y = torch.empty(R, D, dtype=f32, device=dev)
        d_y16 = torch.empty(R, D, dtype=bf16, device=dev)
        dg, db = torch.zeros(1, D, dtype=f32, device=dev), torch.zeros(1, D, dtype=f32, device=dev)
        ops.layernorm_bwd(y, None, g, d_out.detach().reshape(R, D).float().contiguous(), ctx.eps, R, D, d_x=d_y, d_x16=d_y16,
                          d_gamma=dg, d_beta=db)
        d_b = _colsum(d_y)
        d_yT, _ = _tcast(d_y, R, D)
        xT, _ = _tcast(x16, R, kp)
        d_w = _wgrad(d_yT, xT, D, kp)[:, :din]
        d_x = None
        if ctx.needs_input_grad[0]:
            d_x2 = torch.empty(R, kp, dtype=f32, device=dev)
            ops.linear(d_y16, w16.t().contiguous(), d_x2, M=R, N=kp, K=D)
            d_x = d_x2[:, :din].reshape(*ctx.lead, din).to(ctx.x_dtype)
        return d_x, None, d_w, d_b, dg[0], db[0], None


def linear_ln_train(x: Optional[torch.Tensor], lin: nn.Linear, ln: nn.LayerNorm, x16: Optional[torch.Tensor] = None):
    """LayerNorm(Linear(x)) with gradients; pass `x16` (bf16 [R, pad64(d_in)], e.g. Fourier features) instead of x when
    the input needs no gradient."""
    return _LinearLNFn.apply(x, x16, lin.weight, lin.bias, ln.weight, ln.bias, ln.eps)


# ------------------------------------------------------------------------------------------------------------
# Linear - ReLU - LayerNorm - Dropout - Linear
# ------------------------------------------------------------------------------------------------------------
class MlpHeadTrain:
    """One forward pass' operand copies of a get_mlp_head module + per-call forward / backward.  `neg_inf_cols`: output
    columns overwritten by -inf after the head (MaskHeadSegLevel's filter_out_classes): folded into the last bias, and
    their gradient is dropped.  Dropout (module in .train(), p > 0) uses the counter RNG: `seed` device int32, `site`."""

    def __init__(self, seq: nn.Sequential, neg_inf_cols=None, seed=None):
        l0, ln, l4 = seq[0], seq[2], seq[4]
        dev = l0.weight.device
        self.dev = dev
        self.p = float(seq[3].p) if seq.training else 0.0
        self.seed = seed
        if self.p > 0.0 and seed is None:
            raise ValueError("MLP head dropout needs the step's device seed")
        f = lambda t: t.detach().float().contiguous()                 # noqa: E731
        w16 = lambda t: t.detach().to(bf16).contiguous()              # noqa: E731
        self.Din, self.Hd, self.C = l0.in_features, l0.out_features, l4.out_features
        self.Cp = pad64(self.C)
        b4 = f(l4.bias).clone()
        self.cols = neg_inf_cols
        if neg_inf_cols is not None:
            b4[..., neg_inf_cols] = float("-inf")
        w4t = torch.zeros(self.Hd, self.Cp, dtype=bf16, device=dev)
        w4t[:, :self.C] = l4.weight.detach().t()
        self.w = dict(w0=w16(l0.weight), w0t=w16(l0.weight.t()), b0=f(l0.bias), g=f(ln.weight)[None], be=f(ln.bias)[None],
                      eps=ln.eps, w4=w16(l4.weight), w4t=w4t, b4=b4)
        self.grads: Dict[str, torch.Tensor] = {}

    def _acc(self, name, g):
        self.grads[name] = g if name not in self.grads else self.grads[name] + g

    def fwd(self, x16: torch.Tensor, R: int, site: int = 0):
        """x16 bf16 [R, Din] -> (out fp32 [R, C], saved)."""
        w, dev, Hd = self.w, self.dev, self.Hd
        h = torch.empty(R, Hd, dtype=f32, device=dev)
        ops.linear(x16, w["w0"], h, M=R, N=Hd, K=self.Din, bias=w["b0"], relu=True)
        hg16 = torch.empty(R, Hd, dtype=bf16, device=dev)                 # ReLU gate of the backward
        ops.cast_bf16(h, hg16)
        hn16 = torch.empty(R, Hd, dtype=bf16, device=dev)
        ops.add_layernorm(h, None, w["g"], w["be"], w["eps"], R, Hd, out_bf16=hn16)
        if self.p > 0.0:
            ops.dropout_bf16(hn16, self.p, self.seed, site)
        out = torch.empty(R, self.C, dtype=f32, device=dev)
        ops.linear(hn16, w["w4"], out, M=R, N=self.C, K=Hd, bias=w["b4"], ldc=self.C)
        return out, dict(x16=x16, h=h, hg16=hg16, hn16=hn16, site=site)

    def bwd(self, sv, d_out: torch.Tensor, xT: Optional[torch.Tensor] = None):
        """d_out [R, C] -> d_x fp32 [R, Din]; parameter gradients accumulate in self.grads ('0.weight', ... '4.bias')."""
        w, dev, Hd, C, Cp = self.w, self.dev, self.Hd, self.C, self.Cp
        R = sv["h"].shape[0]
        dc = torch.zeros(R, Cp, dtype=f32, device=dev)          # output dimension padded to the GEMMs' K granule
        dc[:, :C] = d_out.detach().reshape(R, C)
        if self.cols is not None:
            dc[:, :C][:, self.cols] = 0.0                        # those outputs were overwritten by -inf
        self._acc("4.bias", _colsum(dc)[:C])
        dcT, dc16 = _tcast(dc, R, Cp, want_c=True)
        hnT, _ = _tcast(sv["hn16"], R, Hd)
        self._acc("4.weight", _wgrad(dcT[:C], hnT, C, Hd))
        d_hd16 = torch.empty(R, Hd, dtype=bf16, device=dev)
        ops.linear(dc16, w["w4t"], d_hd16, M=R, N=Hd, K=Cp)
        if self.p > 0.0:
            ops.dropout_bf16(d_hd16, self.p, self.seed, sv["site"])
        d_hn = d_hd16.float()
        d_h = torch.empty(R, Hd, dtype=f32, device=dev)
        dg, db = torch.zeros(1, Hd, dtype=f32, device=dev), torch.zeros(1, Hd, dtype=f32, device=dev)
        ops.layernorm_bwd(sv["h"], None, w["g"], d_hn, w["eps"], R, Hd, d_x=d_h, d_gamma=dg, d_beta=db)
        self._acc("2.weight", dg[0])
        self._acc("2.bias", db[0])
        self._acc("0.bias", _colsum(d_h, gate=sv["hg16"]))
        d_preT, d_pre16 = _tcast(d_h, R, Hd, want_c=True, gate=sv["hg16"])
        if xT is None:
            xT, _ = _tcast(sv["x16"], R, self.Din)
        self._acc("0.weight", _wgrad(d_preT, xT, Hd, self.Din))
        d_x = torch.empty(R, self.Din, dtype=f32, device=dev)
        ops.linear(d_pre16, w["w0t"], d_x, M=R, N=self.Din, K=Hd)
        return d_x


class _MlpHeadFn(torch.autograd.Function):
    """forward(seq, x (..., Din), *seq parameters) -> (..., C) fp32."""

    @staticmethod
    def forward(ctx, seq, x, *params):
        dev = x.device
        seed = None
        if seq.training and float(seq[3].p) > 0.0:
            seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int32, device=dev)
        t = MlpHeadTrain(seq, None, seed)
        lead = x.shape[:-1]
        R = x.numel() // t.Din
        x16 = torch.empty(R, t.Din, dtype=bf16, device=dev)
        ops.cast_bf16(x.detach().reshape(R, t.Din).float().contiguous(), x16)
        out, sv = t.fwd(x16, R)
        ctx.t, ctx.sv, ctx.seq, ctx.lead, ctx.x_dtype = t, sv, seq, lead, x.dtype
        return out.view(*lead, t.C)

    @staticmethod
    def backward(ctx, d_out):
        t, seq = ctx.t, ctx.seq
        d_x = t.bwd(ctx.sv, d_out)
        grads = []
        for name, p in seq.named_parameters():
            g = t.grads.get(name)
            grads.append(None if g is None else g.reshape(p.shape).to(p.dtype))
        ctx.t = ctx.sv = None
        return (None, d_x.view(*ctx.lead, t.Din).to(ctx.x_dtype) if ctx.needs_input_grad[1] else None, *grads)


def mlp_head_train(seq: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    return _MlpHeadFn.apply(seq, x, *list(seq.parameters()))
