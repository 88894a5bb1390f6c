"""B200-native (sm_100a) implementation of IMP's prototype-fusion hot path.

Host side: Python/PyTorch (device memory, streams, autograd glue, torch.distributed); compute:
hand-written CUDA in ``csrc/`` behind the C-ABI of ``include/imp_hotpath.h``.  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "kernels", "ops", "modularity", "omics", "model", "prototypes", "step", "parallel"]
