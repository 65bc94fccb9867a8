"""mmdti_b200 — B200-native (sm_100a) drop-in for the MM-DTI training hot path.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/mmdti_b200.h)
  _lib.py    ctypes binding (fails loudly when the library is missing: no fallback)
  ops.py     torch.autograd.Function wrappers around the C ABI
  models/    host-side mirror of the reference's module interface
             (models/transformers.py, encoder.py, infonce.py, contrastive.py, loss.py, fds.py)
  dist.py    data-parallel plumbing (all-gather of contrastive operands, gradient all-reduce)
"""
from . import _lib                                    # noqa: F401
from .config import precision, set_precision         # noqa: F401

__all__ = ["precision", "set_precision"]
