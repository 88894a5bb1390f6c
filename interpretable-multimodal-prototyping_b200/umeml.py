"""Drop-in ``umeml`` (the non-GAN variant, medmm/modeling/models/umeml.py:83-221) on the same sm_100a kernels.

Differences from ``umeml_gan`` that matter here: the bags are NOT sentinel-stripped (``path_net`` runs on the whole
``img`` tensor, :160-168), ``p_proto`` is an ``nn.Parameter`` that receives gradients (:139), the omics enter as two
tokens -- ``omic_net`` and ``g_omic_net`` applied to the whole gene vector (:166-171, so ``DATASET.OMIC.DIM`` must be
1000) --, the bottleneck block is a plain concatenation (:70-80), and training returns ``(logits, modular_loss)``.
The reference broadcasts its ``(1,1,D)`` encoder tokens with ``torch.concat`` and indexes ``t_path[0]`` (:184-207), so it
only runs with one slide per batch; this module keeps that contract but works for any batch size (row j of the output
is what the reference returns for slide j alone)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from . import modularity as _mod
from . import ops
from .registry import MODEL_REGISTRY
from .token_tail import Block, TransLayer
from .umeml_gan import _cfg


class ConcatBottleneckBlock(nn.Module):
    """medmm/modeling/models/umeml.py:54-80 (class ``BottleneckAttentionBlock`` there)."""

    def __init__(self, dim: int = 256, n_reg: int = 2):
        super().__init__()
        self.bottle_tokens = nn.Parameter(torch.empty(1, n_reg, dim).uniform_())
        self.encoders = nn.ModuleList([Block(dim=dim) for _ in range(2)])

    def forward(self, x_path, x_omic):
        path_len, token_len = x_path.shape[1], self.bottle_tokens.shape[1]
        x = torch.cat([x_path, self.bottle_tokens.expand(x_path.shape[0], -1, -1), x_omic], dim=1)
        for blk in self.encoders:
            x = blk(x)
        return (x[:, :1], x[:, 1:path_len], x[:, path_len + token_len: path_len + token_len + 1], x[:, path_len + token_len + 1:])


class UMEML(nn.Module):
    def __init__(self, cfg, num_classes, omic_sizes=None):
        super().__init__()
        hidden = int(_cfg(cfg, "MODEL.HIDDEN_DIM", 256))
        if hidden != ops.D:
            raise NotImplementedError("kernels are built for MODEL.HIDDEN_DIM = 256")
        self.cfg = cfg
        self.dropout = float(_cfg(cfg, "MODEL.DROPOUT", 0.25))
        self.omic_input_dim = int(_cfg(cfg, "DATASET.OMIC.DIM", 1000))
        self.fusion = _cfg(cfg, "MODEL.FUSION", "concat")
        self.n_proto = int(_cfg(cfg, "MODEL.UMEML.PROTOTYPES", 6))
        self.n_reg = int(_cfg(cfg, "MODEL.UMEML.REGISTERS", 3))
        self.path_net = nn.Sequential(nn.Linear(int(_cfg(cfg, "DATASET.PATH.DIM", 512)), hidden), nn.ReLU(), nn.Dropout(self.dropout))
        self.omic_net = nn.Sequential(nn.Linear(self.omic_input_dim, hidden), nn.ReLU(), nn.Dropout(self.dropout))
        self.g_omic_net = nn.Sequential(nn.Linear(1000, hidden), nn.ReLU(), nn.Dropout(self.dropout))
        self.proto_g_blocks = nn.ModuleList([ops.PathProtoGenerator(dim=hidden) for _ in range(2)])
        self.omic_encoder = nn.Sequential(*[Block(dim=hidden) for _ in range(2)])
        self.layer_norm_p = nn.LayerNorm(hidden)
        self.layer_norm_o = nn.LayerNorm(hidden)
        self.path_decoder = TransLayer(dim=hidden)
        self.omic_decoder = TransLayer(dim=hidden)
        self.bottleattn = ConcatBottleneckBlock(dim=hidden, n_reg=self.n_reg)
        self.p_proto = nn.Parameter(ops.reset_prototypes(self.n_proto, hidden))
        self.p_encoder_token = nn.Parameter(torch.empty(1, 1, hidden).uniform_())
        self.o_encoder_token = nn.Parameter(torch.empty(1, 1, hidden).uniform_())
        if self.fusion != "concat":
            raise NotImplementedError("MODEL.FUSION must be 'concat' (the shipped survival configs)")
        self.mm = nn.Sequential(nn.Linear(hidden * 2, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU())
        self.classifier = nn.Linear(hidden, num_classes)

    def forward(self, batch: Dict):
        img, x_omic = batch["img"], batch["omic"]
        bsz, n, d = img.shape
        # every row of img is a patch here: no sentinel strip in this model (umeml.py:160-168)
        x = ops._to_bf16(img.reshape(bsz * n, d).float())
        cu = torch.arange(0, (bsz + 1) * n, n, dtype=torch.int32, device=img.device)
        p = self.dropout if self.training else 0.0
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) & 0x7FFFFFFF if p > 0 else 0
        p_proto, h = ops.proto_fusion(x, cu, n, self.p_proto, self.path_net[0].weight, self.path_net[0].bias,
                                      [ops.block_params(b) for b in self.proto_g_blocks], p_drop=p, seed=seed)
        xo = x_omic.reshape(bsz, 1, -1).float()
        h_omic_bag = torch.cat([self.omic_net(xo), self.g_omic_net(xo.detach())], dim=1)                  # (B,2,256)
        h_omic = self.omic_encoder(torch.cat([self.o_encoder_token.expand(bsz, -1, -1), h_omic_bag], dim=1))
        h_path = self.path_decoder(torch.cat([self.p_encoder_token.expand(bsz, -1, -1), p_proto], dim=1))
        h_omic = self.layer_norm_o(self.omic_decoder(h_omic))
        h_path = self.layer_norm_p(h_path)
        t_path, _, t_omic, _ = self.bottleattn(h_path, h_omic)
        logits = self.classifier(self.mm(torch.cat([t_path, t_omic], dim=2)).reshape(bsz, -1))
        if not self.training:
            return logits
        terms = _mod.modularity_terms(h, cu, n, p_proto, h_omic)                                          # umeml.py:195-198
        return logits, terms[:, 0].sum() + terms[:, 1].sum()


@MODEL_REGISTRY.register()
def umeml(**kwargs):
    return UMEML(**kwargs)
