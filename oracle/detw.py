"""Deterministic, RNG-independent tensors for fixtures whose weights are too large to
commit (TEST INFRASTRUCTURE).  splitmix64 over the flat index -> uniform with the
requested standard deviation; bit-stable across torch/numpy versions."""
import numpy as np
import torch

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def det_tensor(shape, seed, std=0.02, mean=0.0):
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        x = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)          # [0,1)
    v = (u - 0.5) * (2.0 * np.sqrt(3.0) * std) + mean
    return torch.from_numpy(v.astype(np.float32)).reshape(shape)


def det_state_dict(shapes, seed=1, std=0.02):
    """shapes: {name: shape}.  LayerNorm weights ~ 1 + U, everything else ~ U(std)."""
    out = {}
    for i, (k, shp) in enumerate(sorted(shapes.items())):
        if "layer_norm.weight" in k or "LayerNorm.weight" in k:
            out[k] = det_tensor(shp, seed * 1000 + i, 0.1, 1.0)
        elif k.endswith("bias") and "gbf.bias" not in k:
            out[k] = det_tensor(shp, seed * 1000 + i, 0.05)
        else:
            out[k] = det_tensor(shp, seed * 1000 + i, std)
    if "embed_tokens.weight" in out:
        out["embed_tokens.weight"][0].zero_()
    return out
