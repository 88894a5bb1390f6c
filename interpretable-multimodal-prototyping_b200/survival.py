"""Survival loss and risk score on the device: the last host round trips of a training / test
step.  Same signatures as the reference functions they restate.

  * ``nll_loss_new(logits, Y, c, alpha=0.0, eps=1e-7, reduction='mean')``  medmm/loss/loss.py:28-95 -- discrete-time
    survival negative log-likelihood; ``logits`` is the model's output TUPLE (the reference indexes ``logits[0]``).
  * ``risk(logits)``   medmm/evaluation/evaluator.py:369-382 -- risk = -sum_k prod_{j<=k} (1 - sigmoid(logit_j));
    ``survival_curve`` returns the per-bin survival the evaluator also stores.
Both are a handful of fused elementwise ops on (B, num_bins) tensors and stay on the device the logits live on."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _survival(x: torch.Tensor) -> torch.Tensor:
    """prod_{j<=k} (1 - sigmoid(x_j)) = exp(-sum_{j<=k} softplus(x_j)).  Same value as the reference's
    ``torch.cumprod(1 - hazards, dim=1)``; written as a cumulative sum because cumprod's backward probes its input for
    zeros with a device->host sync (``(input == 0).any().item()``), which stalls the stream and cannot be captured in a
    CUDA graph."""
    return torch.exp(-torch.cumsum(F.softplus(x), dim=1))


def nll_loss_new(logits, Y, c, alpha=0.0, eps=1e-7, reduction="mean"):
    x = logits[0] if isinstance(logits, (tuple, list)) else logits
    bsz = x.shape[0]
    y = Y.to(device=x.device, dtype=torch.int64).view(bsz, 1)
    cens = c.to(device=x.device, dtype=torch.int64).view(bsz, 1)
    hazards = torch.sigmoid(x)
    surv = _survival(x)
    surv_pad = torch.cat([torch.ones_like(hazards[:, :1]), surv], dim=1)          # S(-1) = 1
    s_prev = surv_pad.gather(1, y).clamp(min=eps)
    h_this = hazards.gather(1, y).clamp(min=eps)
    s_this = surv_pad.gather(1, y + 1).clamp(min=eps)
    uncensored = -(1 - cens) * (torch.log(s_prev) + torch.log(h_this))
    censored = -cens * torch.log(s_this)
    loss = censored + uncensored
    if alpha is not None:
        loss = (1 - alpha) * loss + alpha * uncensored
    if reduction == "mean":
        return loss.mean()
    if reduction == "sum":
        return loss.sum()
    raise ValueError("Bad input for reduction: {}".format(reduction))


def survival_curve(logits: torch.Tensor) -> torch.Tensor:
    return _survival(logits)


def risk(logits: torch.Tensor) -> torch.Tensor:
    return -survival_curve(logits).sum(dim=1)
