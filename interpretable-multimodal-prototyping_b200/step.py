"""One training step of the hot path (what bench.py, smoke() and the multi-GPU tests drive).

The reference step is MBTRAIN.forward_backward (medmm/engine/mbtrain.py:108-267): forward of
UMEML_GAN, loss = NLL + KD + modularity, backward, Adam.  The hot path ends at the prototype /
omic tokens; the token-level tail is represented here by fixed cotangents on
those tokens, so every gradient that crosses the hot path is produced: dW1, db1, both prototype
blocks, the six omic encoders, and the modularity gradients into both token groups."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .model import IMPHotPath


class HotPathStep(nn.Module):
    """model + the 1-token omic prefix the reference prepends before the modularity call
    (o_encoder_token, umeml_gan.py:319-320,443-447) so the omic group has K+1 = 7 tokens."""

    def __init__(self, model: IMPHotPath, with_modularity: bool = True):
        super().__init__()
        self.model = model
        self.with_modularity = with_modularity
        self.o_encoder_token = nn.Parameter(torch.rand(1, 1, 256))

    def forward(self, batch: Dict, cot_proto: torch.Tensor, cot_omic: Optional[torch.Tensor], lengths=None):
        out = self.model(batch, lengths)
        loss = (out["p_proto"] * cot_proto).sum()
        h_omic = None
        if out["h_omic_bag"] is not None:
            bsz = out["h_omic_bag"].shape[0]
            h_omic = torch.cat([self.o_encoder_token.expand(bsz, -1, -1), out["h_omic_bag"]], dim=1)
            if cot_omic is not None:
                loss = loss + (h_omic * cot_omic).sum()
        if self.with_modularity and self.training:
            loss = loss + self.model.modularity_loss(out, out["p_proto"], h_omic)
        return loss


def allreduce_gradients(module: nn.Module, world_size: int) -> None:
    """Mean of the per-rank gradients in one flat bucket (NCCL over NVLink; gloo in CPU tests)."""
    import torch.distributed as dist
    params = [p for p in module.parameters() if p.grad is not None]
    if not params or world_size == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat)
    flat.div_(world_size)
    o = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
