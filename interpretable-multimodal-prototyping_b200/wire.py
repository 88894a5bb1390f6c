"""Host-side wire format of a batch of slides.

The reference loader (medmm/data/data_manager.py:347-416) turns every slide into a dense fp32 tensor padded to
10 000 rows with -10000 (``pad_bag``, :356-367,387) and the default collate stacks them: 20.5 MB per slide cross
PCIe, most of it padding for small bags, and the model then strips the padding again with a host sync per slide
(umeml_gan.py:401-410).  At the speed of the sm_100a kernels that copy is the end-to-end limit (8 GPUs x 1.07 GB
per step from one host memory system), so the native wire format is:

    x_packed     (R, 512) bf16   valid rows of all bags back to back, pinned host memory
    cu_seqlens   (B+1,)   int32  row offsets
    max_len      int             host-side bound for grid sizing
    omic / mol   (B, G)   fp32   unchanged

bf16 is what the kernels consume anyway (``ops.strip_and_pack`` rounds to bf16 on the device with the same
round-to-nearest-even), so results are bit-identical to the padded fp32 path; bytes per slide drop from
``10000*512*4`` to ``N*512*2``.  ``collate_packed`` is a drop-in ``collate_fn`` for the reference's DataLoader
(samples are the dicts ``DatasetWrapper_UMEML.__getitem__`` returns, :395-403); the padded fp32 layout stays
supported by the model as the compatibility path (``batch["img"]``)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch

SENTINEL = -10000.0          # data_manager.py:387
PAD_ROWS = 10000             # data_manager.py:387


def pad_bag(bag: torch.Tensor, target_rows: int = PAD_ROWS) -> torch.Tensor:
    """The reference layout of one slide (data_manager.py:356-367): rows past the bag filled with -10000; a bag
    with at least ``target_rows`` rows is returned unpadded (and therefore without sentinel)."""
    bag = bag.float()
    n = bag.shape[0]
    if n >= target_rows:
        return bag
    out = torch.full((target_rows, bag.shape[1]), SENTINEL, dtype=torch.float32)
    out[:n] = bag
    return out


def bag_rows(img: torch.Tensor) -> int:
    """Number of real rows of a padded slide: index of the first row holding a -10000 element, else all rows
    (umeml_gan.py:404-409; the no-sentinel case follows ops.strip_and_pack)."""
    hit = (img == SENTINEL).any(dim=1)
    idx = torch.nonzero(hit)
    return int(idx[0, 0]) if idx.numel() else int(img.shape[0])


def pack_bags(bags: Sequence[torch.Tensor], pin: bool = True, strip: bool = True) -> Dict[str, object]:
    """bags: per-slide (N_i, 512) feature matrices (fp32/bf16; padded ones are stripped when ``strip``) ->
    {"x_packed", "cu_seqlens", "max_len"} with x_packed in (pinned) host memory."""
    lens: List[int] = []
    for b in bags:
        lens.append(bag_rows(b) if (strip and b.dtype != torch.bfloat16) else int(b.shape[0]))
    total = sum(lens)
    width = int(bags[0].shape[1]) if len(bags) else 512
    x = torch.empty((total, width), dtype=torch.bfloat16)
    if pin and torch.cuda.is_available():
        x = x.pin_memory()
    o = 0
    for b, n in zip(bags, lens):
        x[o:o + n].copy_(b[:n])                      # fp32 -> bf16 round-to-nearest-even, same as the device cast
        o += n
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    if lens:
        cu[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0).to(torch.int32)
    return {"x_packed": x, "cu_seqlens": cu, "max_len": max(lens) if lens else 0}


def collate_packed(samples: Sequence[Dict]) -> Dict[str, object]:
    """``collate_fn`` for ``DatasetWrapper_UMEML`` samples: same keys as the default collate of the reference
    (label, survival_month, censorship, mol, patient_id, index) except that ``img`` (B,10000,512) fp32 is replaced by
    the packed triple.  ``mol`` is also exposed as ``omic`` (what MBTRAIN passes to the model, mbtrain.py:152)."""
    out: Dict[str, object] = pack_bags([s["img"] for s in samples])
    for key in ("label", "survival_month", "censorship", "mol"):
        if key in samples[0]:
            out[key] = torch.stack([torch.as_tensor(s[key]) for s in samples])
    if "mol" in out:
        out["omic"] = out["mol"]
    for key in ("patient_id", "index"):
        if key in samples[0]:
            out[key] = [s[key] for s in samples]
    return out


def to_device(batch: Dict[str, object], device, non_blocking: bool = True) -> Dict[str, object]:
    """H2D of every tensor of a collated batch (asynchronous for pinned sources)."""
    return {k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def unpack_to_reference_layout(batch: Dict[str, object], target_rows: int = PAD_ROWS) -> torch.Tensor:
    """Inverse for tests / the compatibility path: packed triple -> (B, target_rows, 512) fp32 with -10000 padding."""
    x, cu = batch["x_packed"], batch["cu_seqlens"]
    nb = cu.numel() - 1
    rows = max(target_rows, int(batch["max_len"]))
    img = torch.full((nb, rows, x.shape[1]), SENTINEL, dtype=torch.float32)
    for i in range(nb):
        a, b = int(cu[i]), int(cu[i + 1])
        img[i, :b - a] = x[a:b].float()
    return img


class RunningOmicMeans:
    """Cohort means of the gene vector without concatenating the whole training set (the reference makes one full
    pass over the loader and concatenates every ``batch["mol"]``, trainer.py:286-291)."""

    def __init__(self):
        self.total: Optional[torch.Tensor] = None
        self.count = 0

    def update(self, mol: torch.Tensor) -> None:
        s = mol.double().sum(dim=0)
        self.total = s if self.total is None else self.total + s
        self.count += int(mol.shape[0])

    def result(self) -> torch.Tensor:
        if self.total is None:
            raise ValueError("no batches seen")
        return (self.total / self.count).float()


def omic_means(batches: Iterable[Dict[str, object]]) -> torch.Tensor:
    acc = RunningOmicMeans()
    for b in batches:
        acc.update(b["mol"])
    return acc.result()
