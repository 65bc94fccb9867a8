"""Fused multi-tensor Adam on the C ABI (mmdti_adam_step): the reference's optimizer (tasks/trainer.py:160-162,
``Adam(lr, eps=1e-6)``) as ONE kernel launch per step that also refreshes the bf16 shadows of the encoder's GEMM
weights, so the separate fp32 -> bf16 cast pass of every step disappears.  CUDA-graph safe: the step count lives on
the device and the kernel arguments are a device-resident pointer table."""
import ctypes

import torch

from . import _lib
from ._lib import call, i32, stream_ptr


class FusedAdam:
    """Same update rule and state meaning as ``torch.optim.Adam(params, lr, betas, eps)`` (no weight decay / amsgrad).

    ``shadows``: optional {parameter: bf16 tensor of the same shape}; every step writes the updated parameter into its
    shadow (see ``TransformerEncoderWithPair.use_external_lowp``).  ``grad_scale`` multiplies every gradient first.
    ``grad_source``: optional callable parameter -> tensor read INSTEAD of ``p.grad`` (e.g. the flat all-reduced
    buffers of ``dist.OverlappedGradReducer(keep_flat=True)``, which saves the copy back into ``p.grad``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, shadows=None, grad_scale=1.0, grad_source=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.MMDTIError("FusedAdam needs contiguous fp32 CUDA parameters")
        self.lr, self.betas, self.eps, self.grad_scale = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(grad_scale)
        dev = self.params[0].device
        self.device = dev
        total = sum(p.numel() for p in self.params)
        # 16-byte aligned slices of two flat state buffers
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.exp_avg = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=dev, dtype=torch.float32)
        self._offs = offs
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int64)
        self.shadows = dict(shadows or {})
        chunk = _lib.lib().mmdti_adam_chunk()
        ch = []
        for ti, p in enumerate(self.params):
            ch += [(ti, s) for s in range(0, p.numel(), chunk)]
        self.chunks = torch.tensor(ch, dtype=torch.int32, device=dev).contiguous()
        self.nchunks = len(ch)
        self._host_table = torch.empty((len(self.params), 6), dtype=torch.int64).pin_memory()
        self.table = torch.empty((len(self.params), 6), dtype=torch.int64, device=dev)
        self._grad_ptrs = None
        self.grad_source = grad_source
        self.total_elements = total

    def state_for(self, p):
        """(exp_avg, exp_avg_sq) views of one parameter (same meaning as torch.optim.Adam's state)."""
        i = next(k for k, q in enumerate(self.params) if q is p)
        o, n = self._offs[i], p.numel()
        return self.exp_avg[o:o + n].view_as(p), self.exp_avg_sq[o:o + n].view_as(p)

    def _refresh_table(self):
        src = self.grad_source if self.grad_source is not None else (lambda q: q.grad)
        ptrs = tuple(src(p).data_ptr() for p in self.params)
        if ptrs == self._grad_ptrs:
            return
        t = self._host_table
        for i, p in enumerate(self.params):
            g = src(p)
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise _lib.MMDTIError("FusedAdam needs contiguous fp32 gradients")
            o = self._offs[i]
            sh = self.shadows.get(p)
            t[i, 0], t[i, 1] = p.data_ptr(), g.data_ptr()
            t[i, 2], t[i, 3] = self.exp_avg.data_ptr() + 4 * o, self.exp_avg_sq.data_ptr() + 4 * o
            t[i, 4] = sh.data_ptr() if sh is not None else 0
            t[i, 5] = p.numel()
        self.table.copy_(t, non_blocking=True)
        self._grad_ptrs = ptrs

    @torch.no_grad()
    def step(self):
        if self.grad_source is None:
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
        self._refresh_table()
        self.step_count.add_(1)
        d = ctypes.c_double
        call("mmdti_adam_step", self.table, self.chunks, i32(self.nchunks), self.step_count, d(self.lr), d(self.betas[0]),
             d(self.betas[1]), d(self.eps), d(self.grad_scale), stream_ptr())

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
