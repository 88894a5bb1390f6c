"""compute_modularity (medmm/modeling/ops/utils.py:205-228) behind the sm_100a kernels.

The reference call is ``compute_modularity(c (1,P,D), x (1,N,D), temp=0.1, grid=False) -> scalar``
with x detached (utils.py:208).  Here the same signature is kept, plus a batched varlen form
that evaluates the prototype tokens and the omic tokens of every bag in one sweep of the patch
graph (the reference runs two separate calls per slide, umeml_gan.py:520-521)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import kernels
from .ops import _to_bf16


def normalize_tokens(c: torch.Tensor) -> torch.Tensor:
    """c (B,P,D) -> c / ||c||_2 taken ACROSS TOKENS per feature: the reference passes ``c.T`` of a
    3-D tensor to F.normalize(dim=1) (utils.py:180,214), which normalises over P, eps 1e-12."""
    return c / c.norm(dim=1, keepdim=True).clamp_min(1e-12)


class _ModularityFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, cu, max_len, chat, n1, n2, temp):
        loss, dchat = kernels.modularity(h, cu, max_len, chat.contiguous(), n1, n2, temp)
        ctx.save_for_backward(dchat)
        ctx.n1, ctx.n2 = n1, n2
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dchat,) = ctx.saved_tensors
        n1 = ctx.n1
        g = torch.cat([dloss[:, 0:1].expand(-1, n1), dloss[:, 1:2].expand(-1, ctx.n2)], dim=1)   # (B, Pt)
        return None, None, None, dchat * g.unsqueeze(-1), None, None, None


class _ShardedModularityFn(torch.autograd.Function):
    """One bag sharded by rows over the ranks of ``group``; loss and token gradient are global on every rank."""

    @staticmethod
    def forward(ctx, h_local, row_offset, total_rows, chat, n1, n2, temp, group):
        loss, dchat = kernels.modularity_sharded(h_local, row_offset, total_rows, chat.contiguous(), n1, n2, temp, group)
        ctx.save_for_backward(dchat)
        ctx.n1, ctx.n2 = n1, n2
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (dchat,) = ctx.saved_tensors
        g = torch.cat([dloss[:, 0:1].expand(-1, ctx.n1), dloss[:, 1:2].expand(-1, ctx.n2)], dim=1)
        return None, None, None, dchat * g.unsqueeze(-1), None, None, None, None


def modularity_terms_sharded(h_local: torch.Tensor, row_offset: int, total_rows: int, c_proto: torch.Tensor,
                             c_omic: Optional[torch.Tensor] = None, temp: float = 0.1, group=None) -> torch.Tensor:
    """Giant-bag mode: ONE bag of ``total_rows`` patches, this rank holding the rows
    [row_offset, row_offset + len(h_local)) (``parallel.shard_bounds`` produces such windows); c_proto (1,P,256)
    and c_omic (1,Q,256) replicated.  -> (1,2) global modularity terms, same on every rank."""
    n1 = c_proto.shape[1]
    chat = normalize_tokens(c_proto.float())
    n2 = 0
    if c_omic is not None:
        n2 = c_omic.shape[1]
        chat = torch.cat([chat, normalize_tokens(c_omic.float())], dim=1)
    return _ShardedModularityFn.apply(h_local.detach(), int(row_offset), int(total_rows), chat, n1, n2, float(temp), group)


def modularity_terms(h: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int, c_proto: torch.Tensor,
                     c_omic: Optional[torch.Tensor] = None, temp: float = 0.1) -> torch.Tensor:
    """h (R,256) bf16 packed (no gradient), c_proto (B,P,256), c_omic (B,Q,256) or None ->
    (B,2): modularity of each token group against each bag (utils.py:205-228)."""
    n1 = c_proto.shape[1]
    chat = normalize_tokens(c_proto.float())
    n2 = 0
    if c_omic is not None:
        n2 = c_omic.shape[1]
        chat = torch.cat([chat, normalize_tokens(c_omic.float())], dim=1)
    return _ModularityFn.apply(h.detach(), cu_seqlens, int(max_len), chat, n1, n2, float(temp))


def compute_modularity(c: torch.Tensor, x: torch.Tensor, temp: float = 0.1, grid: bool = False) -> torch.Tensor:
    """Drop-in for medmm.modeling.ops.compute_modularity: c (1,P,D) [grad], x (1,N,D) -> scalar."""
    if grid:
        raise NotImplementedError("grid=True is never used by the reference model (umeml_gan.py:520-521)")
    if c.dim() != 3 or x.dim() != 3 or c.shape[0] != 1 or x.shape[0] != 1:
        raise ValueError("expected c (1,P,D) and x (1,N,D)")
    n = x.shape[1]
    h = _to_bf16(x.detach().reshape(n, x.shape[2]))
    cu = torch.tensor([0, n], dtype=torch.int32, device=x.device)
    return modularity_terms(h, cu, n, c)[0, 0]
