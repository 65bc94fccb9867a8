"""Numerical mode of the kernels.

  act   "bf16"  activations / tensor-core operands in bf16, fp32 accumulation (production)
        "fp32"  validation mode: every kernel computes in plain fp32 FMA (1e-5 parity)
  pair  storage type of the (B,H,L,L) pair tensor that is carried through all layers:
        "bf16" | "fp16" | "fp32".  The reference's own CUDA path (fp16 autocast,
        tasks/trainer.py:181-190) keeps it in fp16; fp32 doubles the K2 HBM traffic.
"""
import contextlib

import torch

_STATE = {"act": "bf16", "pair": "bf16"}
_TORCH = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def set_precision(act=None, pair=None):
    if act is not None:
        if act not in ("bf16", "fp32"):
            raise ValueError("act must be 'bf16' or 'fp32'")
        _STATE["act"] = act
        if act == "fp32":
            _STATE["pair"] = "fp32"
    if pair is not None:
        if pair not in _TORCH:
            raise ValueError("pair must be 'bf16', 'fp16' or 'fp32'")
        if _STATE["act"] == "fp32" and pair != "fp32":
            raise ValueError("fp32 validation mode requires an fp32 pair tensor")
        _STATE["pair"] = pair


@contextlib.contextmanager
def precision(act=None, pair=None):
    old = dict(_STATE)
    try:
        set_precision(act, pair)
        yield
    finally:
        _STATE.update(old)


def act_dtype():
    return _TORCH[_STATE["act"]]


def pair_dtype():
    return _TORCH[_STATE["pair"]]


def gpair_dtype():
    """the gradient of the pair tensor is carried in the pair tensor's own dtype."""
    return pair_dtype()


def fp32_mode():
    return _STATE["act"] == "fp32"
