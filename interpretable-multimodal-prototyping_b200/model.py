"""Hot-path members of ``UMEML_GAN`` (medmm/modeling/models/umeml_gan.py:232-687) on the B200 kernels.

``IMPHotPath`` owns exactly the sub-modules the prototype-fusion path touches, under the
reference's attribute names so a reference checkpoint loads with ``load_state_dict(strict=False)``:
``path_net.0.*``, ``omic_net.{k}.0.*``, ``proto_g_blocks.{0,1}.cross_attn.*`` / ``.norm1.*``;
``p_proto`` is a plain tensor as in the reference (:310-315).  Its ``forward`` covers
umeml_gan.py:380-434 (impute, strip, path_net, omic encoders, two prototype blocks) for a whole
batch in a fixed number of launches with no host synchronisation, and ``modularity_loss`` covers
:516-529.  The token-level tail (Nystrom layers, GAN, bottleneck fusion, classifier) is the
"next" row of SURVEY.md 8(f) and stays ordinary PyTorch in the caller.

Unlike the reference the number of path prototypes is independent of the number of
omic groups, so P = 16/32 runs.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import modularity as _mod
from . import omics, ops


def contiguous_groups(sizes: Sequence[int] = tuple(omics.GROUP_SIZES)):
    out, o = [], 0
    for s in sizes:
        out.append(list(range(o, o + s)))
        o += s
    return out


class IMPHotPath(nn.Module):
    def __init__(self, n_proto: int = 6, path_dim: int = 512, hidden_dim: int = 256, dropout: float = 0.25,
                 gene_group_indexes: Optional[Sequence[Sequence[int]]] = None, seed: Optional[int] = None):
        super().__init__()
        if hidden_dim != ops.D:
            raise NotImplementedError("kernels are built for MODEL.HIDDEN_DIM = 256")
        self.n_proto, self.dropout = int(n_proto), float(dropout)
        self.path_net = nn.Sequential(nn.Linear(path_dim, hidden_dim), nn.ReLU(), nn.Dropout(dropout))
        groups = list(gene_group_indexes) if gene_group_indexes is not None else contiguous_groups()
        self.gene_group_indexes = [list(map(int, g)) for g in groups]
        self.omic_net = nn.ModuleList([
            nn.Sequential(nn.Linear(len(g), hidden_dim), nn.ReLU(), nn.Dropout(dropout)) for g in groups])
        self.proto_g_blocks = nn.ModuleList([ops.PathProtoGenerator(dim=hidden_dim) for _ in range(2)])
        flat, offs = [], [0]
        for g in self.gene_group_indexes:
            flat += g
            offs.append(len(flat))
        self.register_buffer("_gene_index", torch.tensor(flat, dtype=torch.int32), persistent=False)
        self._group_offsets = offs
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        self.p_proto = ops.reset_prototypes(self.n_proto, hidden_dim, generator=gen)
        self.omic_means: Optional[torch.Tensor] = None          # trainer sets it (mbtrain.py:284-289)

    def _apply(self, fn, *a, **k):                               # p_proto follows .to()/.cuda() like a buffer
        super()._apply(fn, *a, **k)
        self.p_proto = fn(self.p_proto)
        return self

    @staticmethod
    def _seed() -> int:
        return int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFF

    def encode_omics(self, x_omic: torch.Tensor, insample_without_omic: Optional[torch.Tensor] = None) -> torch.Tensor:
        """umeml_gan.py:380-392,413-419 -> h_omic_bag (B,K,256)."""
        p = self.dropout if self.training else 0.0
        mask = means = None
        if insample_without_omic is not None:
            if self.omic_means is None:
                raise ValueError("omic_means must be set before masked inference (mbtrain.py:284-289)")
            mask = insample_without_omic.to(torch.int32).contiguous()
            means = self.omic_means.to(x_omic.device).float().reshape(-1).contiguous()
        ws = [m[0].weight for m in self.omic_net]
        bs = [m[0].bias for m in self.omic_net]
        return omics._OmicEncodeFn.apply(x_omic.float().contiguous(), self._gene_index, self._group_offsets, mask,
                                         means, p, self._seed() if p > 0 else 0, *ws, *bs)

    def forward(self, batch: Dict, lengths: Optional[Sequence[int]] = None) -> Dict[str, torch.Tensor]:
        """batch: 'img' (B,Npad,512) fp32 in the reference layout (-10000 row padding), or pre-packed
        'x_packed' (R,512) bf16 + 'cu_seqlens' (B+1) int32 + 'max_len'; 'omic' (B,G) or None;
        optional 'insample_without_omic' (B,G).  Returns p_proto (B,P,256), h_omic_bag (B,K,256) or
        None, and the packed patch tokens h (R,256) bf16 with their offsets for the modularity term."""
        if "x_packed" in batch:
            x, cu, max_len = batch["x_packed"], batch["cu_seqlens"], int(batch["max_len"])
        else:
            x, cu, max_len = ops.strip_and_pack(batch["img"], lengths)
        p = self.dropout if self.training else 0.0
        blocks = [ops.block_params(b) for b in self.proto_g_blocks]
        tokens, h = ops.proto_fusion(x, cu, max_len, self.p_proto, self.path_net[0].weight, self.path_net[0].bias,
                                     blocks, p_drop=p, seed=self._seed() if p > 0 else 0)
        h_omic = None
        if batch.get("omic") is not None:
            h_omic = self.encode_omics(batch["omic"], batch.get("insample_without_omic"))
        return {"p_proto": tokens, "h_omic_bag": h_omic, "h": h, "cu_seqlens": cu, "max_len": max_len}

    @staticmethod
    def modularity_loss(out: Dict[str, torch.Tensor], p_proto: torch.Tensor,
                        h_omic: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mean_j mod(p_proto_j, h_j) + mean_j mod(h_omic_j, h_j)   (umeml_gan.py:516-526)."""
        terms = _mod.modularity_terms(out["h"], out["cu_seqlens"], out["max_len"], p_proto, h_omic)
        return terms[:, 0].mean() + (terms[:, 1].mean() if h_omic is not None else 0.0)
