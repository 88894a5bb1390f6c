"""Drop-in ``umeml_gan`` model: the reference's registry name, constructor, ``state_dict`` and ``forward(batch)``
contract (medmm/modeling/models/umeml_gan.py:232-687,704-706) on the B200 hot path.

    build_model("umeml_gan", cfg=cfg, num_classes=4, omic_sizes=1000)        # medmm/engine/mbtrain.py:69-75
    model(batch) -> logits (eval) | 7-tuple (train) | 5-tuple (cca)          # umeml_gan.py:458-459,681-687

What runs where:
  * hot path (rows A0-A8 of SURVEY.md 8): sentinel strip, path_net, the two prototype blocks, omic encoders with
    imputation, the modularity term -- the sm_100a kernels behind ``IMPHotPath`` (whole batch, no host sync);
  * token-level tail (N1): Nystrom layers, bottleneck fusion, fusion MLP, explainers, KD loss, second pass --
    ``token_tail`` (batched torch, device-side pairing instead of 49 ``.item()`` per slide);
  * GAN phase / replace_ratio swap (N2, :461-497): same optimiser steps in the same order.

``cfg`` keys read (all the reference reads, umeml_gan.py:243-262): ``MODEL.UMEML.PROTOTYPES``, ``MODEL.UMEML.REGISTERS``,
``MODEL.DROPOUT``, ``MODEL.HIDDEN_DIM``, ``MODEL.PROJECT_DIM``, ``MODEL.FUSION``, ``MODEL.SIZE``, ``DATASET.PATH.DIM``,
``DATASET.OMIC.DIM``, ``DATASET.ROOT``; additive: ``TRAINER.PREC`` ("fp32" default | "amp" | "bf16": the token tail runs
under bf16 autocast; the N-scaling kernels always feed bf16 operands to the tensor cores and keep fp32 statistics),
``MODEL.UMEML.IMPORTANCE_LOG`` ("sync" default = the reference's appends to ``<set>_path.txt`` inside forward | "async" |
"defer", see ``__init__``).

Differences from the reference, all in directions the reference cannot run:
  * the number of path prototypes is independent of the six gene groups (P = 16 / 32 work; at P = 6 every parameter
    shape equals the reference's, so its checkpoints load with ``strict=True``);
  * ``g_omic_net`` is kept as a parameter holder (checkpoint key) but not evaluated: its output is unused (:421-422)
    and the reference layer cannot even be applied to the gene vector (Linear(1000) on 3354 genes);
  * a bag without sentinel keeps all its rows (the reference reuses a stale variable, :405-410).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels
from . import modularity as _mod
from . import omics
from .model import IMPHotPath, contiguous_groups
from .registry import MODEL_REGISTRY
from .token_tail import (Block, BottleneckAttentionBlock, Discriminator, Generator, TransLayer, transform_importance,
                         transform_importance_to_half_one_point_five)

SIGNATURE_COLUMNS = ["Tumor Suppressor Genes", "Oncogenes", "Protein Kinases", "Cell Differentiation Markers",
                     "Transcription Factors", "Cytokines and Growth Factors"]              # umeml_gan.py:350-355
SIGNATURE_CSV = "DATASET/tcga_glioma/labels/signatures.csv"                               # :348 (relative to the CWD)
MOLECULAR_CSV = "DATASET/tcga_glioma/molecular/TCGA-02-0047-01A-01-BS1.csv"               # :365


def _cfg(cfg, path: str, default=None):
    node = cfg
    for part in path.split("."):
        if node is None:
            return default
        node = node.get(part, None) if isinstance(node, dict) else getattr(node, part, None)
    return default if node is None else node


def gene_group_indexes_from_csv(signature_csv: str = SIGNATURE_CSV, molecular_csv: str = MOLECULAR_CSV) -> Optional[List[List[int]]]:
    """Row indices of the molecular table whose gene_name is in each signature column (umeml_gan.py:347-369);
    None when the two tables are not there (the reference would raise)."""
    if not (os.path.isfile(signature_csv) and os.path.isfile(molecular_csv)):
        return None
    import pandas as pd
    sig, mol = pd.read_csv(signature_csv), pd.read_csv(molecular_csv)
    return [mol.index[mol["gene_name"].isin(sig[col].dropna().tolist())].tolist() for col in SIGNATURE_COLUMNS]


class UMEML_GAN(IMPHotPath):
    def __init__(self, cfg, num_classes, omic_sizes=None):
        hidden = int(_cfg(cfg, "MODEL.HIDDEN_DIM", 256))
        n_proto = int(_cfg(cfg, "MODEL.UMEML.PROTOTYPES", 6))
        groups = _cfg(cfg, "MODEL.UMEML.GENE_GROUP_INDEXES") or gene_group_indexes_from_csv() or contiguous_groups()
        super().__init__(n_proto=n_proto, path_dim=int(_cfg(cfg, "DATASET.PATH.DIM", 512)), hidden_dim=hidden,
                         dropout=float(_cfg(cfg, "MODEL.DROPOUT", 0.25)), gene_group_indexes=groups)
        self.cfg = cfg
        self.root = os.path.abspath(os.path.expanduser(str(_cfg(cfg, "DATASET.ROOT", "."))))
        self.omic_input_dim = int(_cfg(cfg, "DATASET.OMIC.DIM", sum(len(g) for g in groups)))
        self.fusion = _cfg(cfg, "MODEL.FUSION", "concat")
        self.size = _cfg(cfg, "MODEL.SIZE", "small")
        self.n_reg = int(_cfg(cfg, "MODEL.UMEML.REGISTERS", 3))
        self.prec = str(_cfg(cfg, "TRAINER.PREC", "fp32"))
        # "sync": append to <set>_path.txt / <set>_omic.txt inside forward like the reference (default); "async": pinned
        # copy now, file append once the copy has completed; "defer": keep the device tensors of the last forward in
        # ``self.last_importance`` and write nothing (CUDA-graph capture; ``flush_importance_logs`` appends them)
        self.importance_log = str(_cfg(cfg, "MODEL.UMEML.IMPORTANCE_LOG", "async" if _cfg(cfg, "MODEL.UMEML.ASYNC_IMPORTANCE_LOG", False) else "sync"))
        self.last_importance: Dict[str, torch.Tensor] = {}
        n_omic_tok = len(groups) + 1                                   # o_encoder_token + one token per gene group
        n_path_tok = n_proto + 1

        # GAN (umeml_gan.py:243-249); token counts per modality are decoupled (equal at P = 6)
        self.gan_generator_p2o = Generator((n_path_tok, hidden), (n_omic_tok, hidden))
        self.gan_generator_o2p = Generator((n_omic_tok, hidden), (n_path_tok, hidden))
        self.gan_discriminator_o = Discriminator((n_omic_tok, hidden))
        self.gan_discriminator_p = Discriminator((n_path_tok, hidden))
        adam = dict(lr=0.0001, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0001)
        self.gan_opt_gen = torch.optim.Adam(list(self.gan_generator_p2o.parameters()) + list(self.gan_generator_o2p.parameters()), **adam)
        self.gan_opt_dis_o = torch.optim.Adam(self.gan_discriminator_o.parameters(), **adam)
        self.gan_opt_dis_p = torch.optim.Adam(self.gan_discriminator_p.parameters(), **adam)

        self.g_omic_net = nn.Sequential(nn.Linear(1000, hidden), nn.ReLU(), nn.Dropout(self.dropout))   # checkpoint key only
        self.omic_encoder = nn.Sequential(*[Block(dim=hidden) for _ in range(2)])
        self.layer_norm_p = nn.LayerNorm(hidden)
        self.layer_norm_o = nn.LayerNorm(hidden)
        self.path_decoder = TransLayer(dim=hidden)
        self.omic_decoder = TransLayer(dim=hidden)
        self.bottleattn = BottleneckAttentionBlock(dim=hidden, n_reg=self.n_reg)
        self.p_encoder_token = nn.Parameter(torch.empty(1, 1, hidden).uniform_())
        self.o_encoder_token = nn.Parameter(torch.empty(1, 1, hidden).uniform_())
        if self.fusion == "concat":
            self.mm = nn.Sequential(nn.Linear(hidden * 2, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU())
        elif self.fusion == "bilinear":
            raise NotImplementedError("MODEL.FUSION='bilinear' (ops/blocks.py:102-184) is outside the hot-path scope; the shipped "
                                      "survival config uses 'concat'")
        else:
            self.mm = None
        self.classifier = nn.Linear(hidden, num_classes)
        self.lambda_cyc = 10
        self.l1_loss = nn.L1Loss()
        self.dis_loss = nn.BCELoss()
        self.explainer_path = nn.Linear(hidden, num_classes, bias=False)
        self.explainer_omic = nn.Linear(hidden, num_classes, bias=False)
        # attributes the trainer mutates (engine/trainer.py:620-626,673-675; mbtrain.py:284-289)
        self.train_gan = False
        self.replace_ratio = 0
        self.cca = False
        self.plot_set = "train"
        self._pending_logs: list = []

    # --------------------------------------------------------------------------------------
    def adversarial_loss(self, D, fake):
        # the reference applies BCEWithLogits to the discriminator's sigmoid output (umeml_gan.py:371-372): kept
        out = D(fake)
        return F.binary_cross_entropy_with_logits(out, torch.ones_like(out))

    def _gan_phase(self, h_path: torch.Tensor, h_omic: torch.Tensor):
        """umeml_gan.py:461-490: one generator step and two discriminator steps inside the forward.  The reference
        back-propagates these three losses through the whole model with retain_graph and lets the trainer's
        zero_grad() (engine/trainer.py:354-357) discard everything but the GAN parameters' steps; detaching the tokens
        here gives the same parameter updates and the same returned values without three extra backward passes over
        the patch bags."""
        hp, ho = h_path.detach(), h_omic.detach()
        fake_omic, fake_path = self.gan_generator_p2o(hp), self.gan_generator_o2p(ho)
        cycle_path, cycle_omic = self.gan_generator_o2p(fake_omic), self.gan_generator_p2o(fake_path)
        gen_loss = (self.adversarial_loss(self.gan_discriminator_o, fake_omic) + self.adversarial_loss(self.gan_discriminator_p, fake_path)
                    + self.lambda_cyc * (self.l1_loss(cycle_omic, ho) + self.l1_loss(cycle_path, hp)))
        self.gan_opt_gen.zero_grad()
        gen_loss.backward()
        self.gan_opt_gen.step()

        def disc_step(D, opt, real, fake):
            pred = torch.cat((D(real), D(fake)), dim=0)
            labels = torch.cat((torch.ones(pred.shape[0] // 2, 1), torch.zeros(pred.shape[0] // 2, 1)), dim=0).to(pred.device)
            loss = self.dis_loss(pred, labels)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss

        dis_p_loss = disc_step(self.gan_discriminator_p, self.gan_opt_dis_p, hp, self.gan_generator_o2p(ho))
        dis_o_loss = disc_step(self.gan_discriminator_o, self.gan_opt_dis_o, ho, self.gan_generator_p2o(hp))
        return gen_loss, dis_p_loss, dis_o_loss

    # --------------------------------------------------------------------------------------
    def _fuse_and_classify(self, h_path, h_omic, patient_id):
        t_path, _, t_omic, _ = self.bottleattn(h_path, h_omic, patient_id)
        if self.mm is None:
            raise ValueError("MODEL.FUSION must be 'concat' (the reference leaves h undefined otherwise, umeml_gan.py:532-543)")
        h = self.mm(torch.cat([t_path, t_omic], dim=2)).reshape(h_path.shape[0], -1)      # :537-541, batched
        return self.classifier(h)

    def _write_importance(self, name: str, rows: torch.Tensor) -> None:
        """Appends one line per sample to ``<plot_set>_<name>.txt`` (umeml_gan.py:576-587).  Asynchronous mode copies the
        rows to pinned memory and writes them when their copy has completed (next forward / ``flush_importance_logs``)."""
        path = self.plot_set + "_" + name + ".txt"
        rows = rows.detach()
        if self.importance_log == "defer":
            self.last_importance[path] = rows
            return
        if self.importance_log == "async" and rows.is_cuda:
            host = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
            host.copy_(rows, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pending_logs.append((path, host, ev))
            return
        with open(path, "a") as f:
            for row in rows.tolist():
                f.write(" ".join(map(str, row)) + "\n")

    def flush_importance_logs(self, wait: bool = True) -> None:
        for path, rows in self.last_importance.items():          # "defer" mode: the rows of the last forward / graph replay
            with open(path, "a") as f:
                for row in rows.tolist():
                    f.write(" ".join(map(str, row)) + "\n")
        keep = []
        for path, host, ev in self._pending_logs:
            if not wait and not ev.query():
                keep.append((path, host, ev))
                continue
            ev.synchronize()
            with open(path, "a") as f:
                for row in host.tolist():
                    f.write(" ".join(map(str, row)) + "\n")
        self._pending_logs = keep

    def _classify_and_explain(self, h_path, h_omic, batch: Dict, T: float):
        """Fusion + classifier, explainers / importance / knowledge distillation, second pass (umeml_gan.py:532-678)."""
        bsz = h_path.shape[0]
        logits = self._fuse_and_classify(h_path, h_omic, batch.get("patient_id"))

        # explainers, importance scores, knowledge distillation (:553-598)
        lp = self.explainer_path(h_path)
        lo = self.explainer_omic(h_omic)
        logits_explained = (lp.mean(dim=1) + lo.mean(dim=1)) / 2
        pred = logits_explained.argmax(dim=1)
        imp_path = torch.gather(lp, 2, pred.view(bsz, 1, 1).expand(bsz, lp.shape[1], 1)).squeeze(-1)
        imp_omic = torch.gather(lo, 2, pred.view(bsz, 1, 1).expand(bsz, lo.shape[1], 1)).squeeze(-1)
        importance_path_ = transform_importance(imp_path)[:, :imp_path.shape[1] - 1]
        importance_omic_ = transform_importance(imp_omic)[:, :imp_omic.shape[1] - 1]
        self._write_importance("path", importance_path_)
        self._write_importance("omic", importance_omic_)
        loss_kd = F.kl_div(F.log_softmax(logits_explained / T, dim=1), F.softmax(logits.detach() / T, dim=1),
                           reduction="batchmean") * (T * T)

        # second pass on importance-weighted tokens (:651-678)
        w_path = transform_importance_to_half_one_point_five(imp_path.detach()).unsqueeze(-1)
        w_omic = transform_importance_to_half_one_point_five(imp_omic.detach()).unsqueeze(-1)
        logits = self._fuse_and_classify(h_path * w_path, h_omic * w_omic, batch.get("patient_id"))
        return logits, loss_kd, importance_path_

    # --------------------------------------------------------------------------------------
    def token_tail(self, p_proto, h_omic_bag, batch: Dict, hot: Optional[Dict] = None, T: float = 5.0):
        """Everything after the hot path (umeml_gan.py:436-687): p_proto (B,P,256), h_omic_bag (B,6,256) or None."""
        bsz = p_proto.shape[0]
        p_proto_before, h_omic_bag_before = p_proto, h_omic_bag
        has_omic = h_omic_bag is not None
        if has_omic:
            h_omic = self.omic_encoder(torch.cat([self.o_encoder_token.expand(bsz, -1, -1), h_omic_bag], dim=1))
        h_path = self.path_decoder(torch.cat([self.p_encoder_token.expand(bsz, -1, -1), p_proto], dim=1))
        h_path = self.layer_norm_p(h_path)
        if has_omic:
            h_omic = self.layer_norm_o(self.omic_decoder(h_omic))
        if self.cca:
            return h_path, (h_omic if has_omic else None), p_proto_before, h_omic_bag_before, "cca"

        gen_loss = dis_p_loss = dis_o_loss = 0
        if self.training and self.train_gan:
            gen_loss, dis_p_loss, dis_o_loss = self._gan_phase(h_path, h_omic)
        if self.training and self.replace_ratio > 0:                    # :492-497, same host RNG stream
            swap = torch.from_numpy(np.random.uniform(0, 1, bsz) > self.replace_ratio).to(h_omic.device)
            h_omic = torch.where(swap.view(-1, 1, 1), self.gan_generator_p2o(h_path), h_omic)

        # missing omics (:499-511).  Masks that are all zero are device-side no-ops instead of host-side branches.
        without = batch.get("without_omic")
        insample = batch.get("insample_without_omic")
        if not has_omic:
            h_omic = self.gan_generator_p2o(h_path)
        elif without is not None or insample is not None:
            h_gen = self.gan_generator_p2o(h_path)
            if h_omic.is_cuda and not torch.is_grad_enabled():
                h_omic, _ = kernels.omic_blend(h_omic.float().contiguous(), h_gen.float().contiguous(),
                                               None if without is None else without.to(torch.int32).contiguous(),
                                               None if insample is None else insample.to(torch.int32).contiguous())
            else:
                if without is not None:
                    h_omic = torch.where((without == 1).view(-1, 1, 1).to(h_omic.device), h_gen, h_omic)
                if insample is not None:
                    r = insample.sum().to(h_omic.dtype) / insample.numel()
                    h_omic = (1 - r) * h_omic + r * h_gen

        # The pair sweep behind this call (33 ms for 32 slides of 16 384 patches) needs p_proto and h_omic only, so it is
        # launched BEFORE fusion, explainers and the second pass: their ~1 500 launches are then issued by the CPU while
        # the sweep runs.  (Running them on a second, high-priority stream next to the sweep inside a captured graph was
        # built and measured: 40.75 -> 40.4 ms per step only -- the sweep owns every SM and its CTAs retire in waves, so
        # the chain of small kernels advances a few launches per 0.6 ms wave; launched eagerly that way it was 110 ms.)
        modular_loss = 0
        if self.training:                                               # :516-526, both token groups in one sweep
            if hot is None:
                raise ValueError("training needs the packed patch tokens of the hot path for the modularity term")
            terms = _mod.modularity_terms(hot["h"], hot["cu_seqlens"], hot["max_len"], p_proto, h_omic)
            modular_loss = terms[:, 0].mean() + terms[:, 1].mean()

        logits, loss_kd, importance_path_ = self._classify_and_explain(h_path, h_omic, batch, T)

        if self.training:
            return logits, modular_loss, gen_loss, dis_p_loss, dis_o_loss, loss_kd, importance_path_
        return logits

    def forward(self, batch: Dict, is_survival: bool = True, T: float = 5.0):
        if self._pending_logs:
            self.flush_importance_logs(wait=False)
        hot = IMPHotPath.forward(self, batch)              # strip, path_net, prototype blocks, omic encoders (+ imputation)
        if self.prec == "bf16" and hot["p_proto"].is_cuda:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return self.token_tail(hot["p_proto"], hot["h_omic_bag"], batch, hot, T)
        return self.token_tail(hot["p_proto"], hot["h_omic_bag"], batch, hot, T)


@MODEL_REGISTRY.register()
def umeml_gan(**kwargs):
    return UMEML_GAN(**kwargs)
