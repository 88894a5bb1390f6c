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


class GraphedStep:
    """The whole forward + backward of a fixed-shape batch captured once in a CUDA graph and replayed: the
    hot path launches ~100 of its own kernels plus ~200 tiny cuBLAS/ATen kernels of the token algebra per step,
    and the host cannot issue them as fast as the GPU runs them once the O(N^2) term is off.

    The input tensors given to ``capture`` are static: refill them in place (``copy_``) between replays.
    Dropout: the seeds passed by value are baked into the graph, so the graph advances a device word that
    every dropout kernel XORs into its seed (``imp_set_seed_offset``) -- each replay draws a new mask.
    Gradients land in ``param.grad`` of the wrapped module (tensors owned by the graph's memory pool)."""

    def __init__(self, runner: Optional[nn.Module]):
        self.runner = runner
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self._seed_word: Optional[torch.Tensor] = None

    def capture(self, batch: Dict, cot_proto: torch.Tensor, cot_omic: Optional[torch.Tensor], lengths=None,
                warmup: int = 3) -> "GraphedStep":
        return self.capture_fn(lambda: self.runner(batch, cot_proto, cot_omic, lengths), [p for p in self.runner.parameters()],
                               cot_proto.device, warmup)

    def capture_fn(self, loss_fn, params, dev, warmup: int = 3) -> "GraphedStep":
        """Capture ``loss = loss_fn(); loss.backward()`` for any fixed-shape, sync-free step (e.g. the whole drop-in
        ``umeml_gan`` model with ``importance_log = "defer"``): same protocol as ``capture``."""
        from . import _lib
        params = list(params)
        self._seed_word = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("imp_set_seed_offset", self._seed_word)
        _lib.profile_enable(False)

        def body():
            for p in params:
                p.grad = None
            self._seed_word.add_(0x61C88647)                      # new keep-masks on every replay
            loss = loss_fn()
            loss.backward()
            return loss

        # the warm-up below runs on a side stream on purpose; AccumulateGrad nodes of earlier eager steps may still
        # be bound to the default stream, which is harmless here
        setw = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if setw is not None:
            setw(False)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                              # warm-up off the capture stream (allocator, attributes)
            for _ in range(warmup):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = body()
        self._params = params
        self._grads = [p.grad for p in params]
        return self

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        for p, g in zip(self._params, self._grads):              # another capture may have re-bound .grad
            p.grad = g
        return self.loss

    def close(self) -> None:
        from . import _lib
        if self._seed_word is not None:          # the library must not keep a pointer into a freed tensor
            with torch.cuda.device(self._seed_word.device):
                _lib.call("imp_set_seed_offset", None)
            self._seed_word = None
        self.graph = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def allreduce_gradients(module: nn.Module, world_size: int) -> None:
    """Mean of the per-rank gradients in one flat bucket (NCCL over NVLink; gloo in CPU tests)."""
    import torch.distributed as dist
    if world_size == 1:
        return
    # the bucket covers EVERY trainable parameter (missing gradients count as zero): the layout is then the same on
    # all ranks even when one rank's batch had no omics and left the omic encoders without gradients
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return
    total = sum(p.numel() for p in params)
    flat = getattr(module, "_imp_grad_bucket", None)
    if flat is None or flat.numel() != total or flat.device != params[0].device:
        flat = torch.empty(total, dtype=torch.float32, device=params[0].device)
        module._imp_grad_bucket = flat
    o = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            flat[o:o + n].zero_()
        else:
            flat[o:o + n].copy_(p.grad.reshape(-1))
        o += n
    dist.all_reduce(flat)
    flat.div_(world_size)
    o = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[o:o + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n
