"""K-means prototype extraction over patch features (BASELINE config 5; NEW functionality --
the reference's extract_prototype_with_plip_train.py holds no k-means)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import kernels


def kmeans_assign(x: torch.Tensor, centroids: torch.Tensor) -> torch.Tensor:
    return kernels.kmeans_assign(x.contiguous(), centroids.contiguous())


def kmeans_fit(x: torch.Tensor, k: int, iters: int = 10, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lloyd iterations on the device: centroids initialised from k distinct rows (seeded), empty
    clusters keep their previous centroid.  Returns (centroids (k,D) fp32, assign (N) int32)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    pick = torch.randperm(x.shape[0], generator=g)[:k].to(x.device)
    mu = x[pick].clone().contiguous()
    assign = kernels.kmeans_assign(x, mu)
    for _ in range(iters):
        sums, counts = kernels.kmeans_update(x, assign, k)
        nz = counts > 0
        mu = torch.where(nz[:, None], sums / counts.clamp_min(1)[:, None].float(), mu).contiguous()
        assign = kernels.kmeans_assign(x, mu)
    return mu, assign
