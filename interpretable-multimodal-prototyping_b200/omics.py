"""Genomic side of the hot path: per-pathway encoders (A7) and missing-omics handling (A8).
Reference: medmm/modeling/models/umeml_gan.py:274-283,380-392,413-419,500-511."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import kernels

GROUP_SIZES = [82, 330, 513, 440, 1538, 451]        # umeml_gan.py:274


class _OmicEncodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gene_index, offsets, mask, means, p_drop, seed, *wb):
        k = len(wb) // 2
        ws, bs = wb[:k], wb[k:]
        out = kernels.omic_encode_fwd(x, gene_index, offsets, [w.detach().contiguous() for w in ws],
                                      [b.detach().contiguous() for b in bs], mask, means, p_drop, seed)
        ctx.save_for_backward(x, gene_index, out, *( [mask, means] if mask is not None else []), *ws)
        ctx.offsets, ctx.p_drop, ctx.k, ctx.has_mask = tuple(offsets), p_drop, k, mask is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        t = ctx.saved_tensors
        x, gene_index, out = t[:3]
        mask, means = (t[3], t[4]) if ctx.has_mask else (None, None)
        ws = t[5:] if ctx.has_mask else t[3:]
        dws, dbs = kernels.omic_encode_bwd(x, gene_index, ctx.offsets, out, dout.contiguous(), ws, mask, means,
                                           ctx.p_drop)
        return (None,) * 7 + tuple(dws) + tuple(dbs)


class OmicEncoders(nn.Module):
    """``omic_net`` of UMEML_GAN: ModuleList of Sequential(Linear(G_k,256), ReLU, Dropout) with the
    reference's state_dict keys ``omic_net.{k}.0.{weight,bias}`` (umeml_gan.py:274-283)."""

    def __init__(self, group_indexes: Sequence[Sequence[int]], hidden: int = 256, dropout: float = 0.25):
        super().__init__()
        self.dropout = float(dropout)
        self.omic_net = nn.ModuleList([
            nn.Sequential(nn.Linear(len(ix), hidden), nn.ReLU(), nn.Dropout(dropout)) for ix in group_indexes])
        flat, offs = [], [0]
        for ix in group_indexes:
            flat += [int(i) for i in ix]
            offs.append(len(flat))
        self.register_buffer("gene_index", torch.tensor(flat, dtype=torch.int32), persistent=False)
        self.group_offsets = offs

    def forward(self, x_omic: torch.Tensor, insample_without_omic: Optional[torch.Tensor] = None,
                omic_means: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x_omic (B,G) -> (B,K,256).  ``insample_without_omic`` (B,G) int: masked genes take the
        training mean (umeml_gan.py:391-392)."""
        p = self.dropout if self.training else 0.0
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if p > 0 else 0
        mask = None
        if insample_without_omic is not None:
            mask = insample_without_omic.to(torch.int32).contiguous()
            if omic_means is None:
                raise ValueError("omic_means is required with insample_without_omic (trainer sets it, mbtrain.py:284-289)")
            omic_means = omic_means.float().contiguous()
        ws = [m[0].weight for m in self.omic_net]
        bs = [m[0].bias for m in self.omic_net]
        return _OmicEncodeFn.apply(x_omic.float().contiguous(), self.gene_index, self.group_offsets, mask,
                                   omic_means if mask is not None else None, p, seed, *ws, *bs)


class _BlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h_omic, h_gen, without_omic, insample_mask):
        out, r = kernels.omic_blend(h_omic.contiguous(), h_gen.contiguous(), without_omic, insample_mask)
        ctx.save_for_backward(r, without_omic if without_omic is not None else r.new_zeros(0))
        ctx.has_wo = without_omic is not None
        return out

    @staticmethod
    def backward(ctx, g):
        r, wo = ctx.saved_tensors
        shape = (-1,) + (1,) * (g.dim() - 1)
        keep = 1.0 - (wo.view(shape) == 1).to(g.dtype) if ctx.has_wo else 1.0
        gh = g * (1.0 - r) * keep
        return gh, g - gh, None, None


def blend_missing_omics(h_omic: Optional[torch.Tensor], h_omic_gen: torch.Tensor,
                        without_omic: Optional[torch.Tensor] = None,
                        insample_without_omic: Optional[torch.Tensor] = None) -> torch.Tensor:
    """umeml_gan.py:500-511 without host syncs: sample-level replacement by the generator output,
    then the batch-ratio blend.  ``h_omic is None`` -> generator output (:506-507)."""
    if h_omic is None:
        return h_omic_gen
    wo = without_omic.to(torch.int32).contiguous() if without_omic is not None else None
    im = insample_without_omic.to(torch.int32).contiguous() if insample_without_omic is not None else None
    if wo is None and im is None:
        return h_omic
    return _BlendFn.apply(h_omic.float(), h_omic_gen.float(), wo, im)
