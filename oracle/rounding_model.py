"""Rounding-point model of the device path (TEST INFRASTRUCTURE ONLY, like the rest of ``oracle/``).

``imp_oracle.hot_path_step`` is the reference arithmetic in full precision.  The sm_100a kernels feed bf16
operands to the tensor cores and keep two N-sized tensors in bf16 (h and dz), so their results differ from
the full-precision oracle by the rounding of exactly these tensors:

    point     tensor                                            where (csrc/)
    "h"       h = relu(x W1^T + b1)            stored bf16      pathnet.cu epilogue
    "q"       q~ (folded queries)              bf16 operand     pool.cu load_gmat_rows / G
    "prob"    exp(S - m) softmax weights       bf16 operand     pool.cu forward probability tile
    "dpool"   dpooled (token cotangent)        bf16 operand     pool.cu G, ops.py delta
    "e"       E = [dS | a]                     bf16 operand     pool.cu dq / dz kernels
    "dz"      dz = dh * [h > 0]                stored bf16      pool.cu dz epilogue 2

This module evaluates the same step as ``hot_path_step`` in fp64 with any subset of those roundings switched
on.  With the empty set it IS the oracle (tests/test_oracle_golden.py checks that on the CPU); with all six it
predicts what the kernels compute, so that (a) the kernels can be held to 1e-3 against it, and (b) the error
each rounding point contributes to every gradient can be attributed (profiles/r02_parity.md).
Reference lines restated: umeml_gan.py:410,425-434 / ops/attention.py:355-533 (see imp_oracle.py)."""
from __future__ import annotations

from typing import Dict, Iterable, Optional, Sequence

import torch

from . import imp_oracle as O

ALL_POINTS = ("h", "q", "prob", "dpool", "e", "dz")


def _r(t: torch.Tensor, on: bool) -> torch.Tensor:
    return t.float().bfloat16().to(t.dtype) if on else t


class _PathNet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, pts):
        z = torch.relu(x @ w1.t() + b1)
        h = _r(z, "h" in pts)
        ctx.save_for_backward(x, h)
        ctx.pts = pts
        return h

    @staticmethod
    def backward(ctx, dh):
        x, h = ctx.saved_tensors
        dz = _r(dh * (h > 0).to(dh.dtype), "dz" in ctx.pts)
        return None, dz.t() @ x, dz.sum(0), None


class _Pool(torch.autograd.Function):
    """pooled_p = sum_n softmax_n(h_n . q~_p) h_n with the kernels' operand roundings."""

    @staticmethod
    def forward(ctx, h, qt, pts):
        qr = _r(qt, "q" in pts)
        s = h @ qr.t()                                    # (N,P)
        m = s.max(dim=0).values
        p = _r(torch.exp(s - m), "prob" in pts)
        l = p.sum(0)
        pooled = (p.t() @ h) / l[:, None]
        lse = m + torch.log(l)
        ctx.save_for_backward(h, qt, pooled, lse)
        ctx.pts = pts
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        h, qt, pooled, lse = ctx.saved_tensors
        pts = ctx.pts
        qr, dpr = _r(qt, "q" in pts), _r(dpooled, "dpool" in pts)
        delta = (dpr * pooled).sum(-1)
        s = h @ qr.t()
        da = h @ dpr.t()
        a = torch.exp(s - lse)
        ds = a * (da - delta)
        dsr, ar = _r(ds, "e" in pts), _r(a, "e" in pts)
        dq = dsr.t() @ h
        dh = dsr @ qr + ar @ dpr
        return dh, dq, None


def hot_path_step_rounded(bags: Sequence[torch.Tensor], params: Dict[str, torch.Tensor], p_proto: torch.Tensor,
                          grad_seed: torch.Tensor, points: Iterable[str] = ALL_POINTS,
                          dtype: torch.dtype = torch.float64):
    """Same contract as ``imp_oracle.hot_path_step(..., with_modularity=False)``: differentiates
    sum(c_final * grad_seed); returns dict(c=(B,P,D), grads={name: tensor}) in ``dtype`` on the inputs' device."""
    pts = frozenset(points)
    unknown = pts - set(ALL_POINTS)
    if unknown:
        raise ValueError("unknown rounding points %s" % sorted(unknown))
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    cs, total = [], 0.0
    for j, x in enumerate(bags):
        h = _PathNet.apply(x.to(dtype), leaves["path_net.0.weight"], leaves["path_net.0.bias"], pts)
        c = p_proto.to(dtype)
        for b in range(2):
            pre = "proto_g_blocks.%d." % b
            in_w, in_b = leaves[pre + "cross_attn.in_proj_weight"], leaves[pre + "cross_attn.in_proj_bias"]
            qt = O.folded_query(c, in_w, in_b)
            pooled = _Pool.apply(h, qt, pts)
            d = c.shape[-1]
            v = pooled @ in_w[2 * d:].t() + in_b[2 * d:]
            o = v @ leaves[pre + "cross_attn.out_proj.weight"].t() + leaves[pre + "cross_attn.out_proj.bias"]
            c = c + O.layer_norm(o, leaves[pre + "norm1.weight"], leaves[pre + "norm1.bias"])
        cs.append(c)
        total = total + (c * grad_seed[j].to(dtype)).sum()
    total.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return {"c": torch.stack([c.detach() for c in cs]), "grads": grads}
