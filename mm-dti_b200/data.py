"""Input-format helpers of the hot path: padding rules (reference utils/util.py:7-105) and
the synthetic-molecule generator used by tests and bench.py (format of
data/conformer.py:204-212: tokens with [CLS]/[SEP], centred coordinates, Euclidean
distance matrix, edge_type = tok_i*|dict| + tok_j; pad token 0, pad distance 0)."""
import torch

from . import _lib
from ._lib import call, i32, i64, stream_ptr


def _pad_size(values, pad_to_length, pad_to_multiple):
    size = max(v.size(0) for v in values)
    size = size if pad_to_length is None else max(size, pad_to_length)
    if pad_to_multiple != 1 and size % pad_to_multiple != 0:
        size = int(((size - 0.1) // pad_to_multiple + 1) * pad_to_multiple)
    return size


def pad_1d_tokens(values, pad_idx, left_pad=False, pad_to_length=None, pad_to_multiple=1):
    size = _pad_size(values, pad_to_length, pad_to_multiple)
    res = values[0].new_full((len(values), size), pad_idx)
    for i, v in enumerate(values):
        (res[i, size - len(v):] if left_pad else res[i, :len(v)]).copy_(v)
    return res


def pad_2d(values, pad_idx, left_pad=False, pad_to_length=None, pad_to_multiple=1):
    size = _pad_size(values, pad_to_length, pad_to_multiple)
    res = values[0].new_full((len(values), size, size), pad_idx)
    for i, v in enumerate(values):
        n = len(v)
        (res[i, size - n:, size - n:] if left_pad else res[i, :n, :n]).copy_(v)
    return res


def pad_coords(values, pad_idx, left_pad=False, pad_to_length=None, pad_to_multiple=1):
    size = _pad_size(values, pad_to_length, pad_to_multiple)
    res = values[0].new_full((len(values), size, 3), pad_idx)
    for i, v in enumerate(values):
        (res[i, size - len(v):, :] if left_pad else res[i, :len(v), :]).copy_(v)
    return res


def synthetic_molecules(B, n_atoms, seed=1234, ragged=False, n_dict=31):
    """B synthetic conformers with up to ``n_atoms`` atoms (+[CLS]/[SEP] => L = n_atoms+2).
    Returns CPU tensors: src_tokens (B,L) int64, src_distance (B,L,L) f32,
    src_edge_type (B,L,L) int64, src_coord (B,L,3) f32."""
    g = torch.Generator().manual_seed(seed)
    L = n_atoms + 2
    if ragged:
        n = torch.randint((n_atoms + 1) // 2, n_atoms + 1, (B,), generator=g)
        n[0] = n_atoms
    else:
        n = torch.full((B,), n_atoms, dtype=torch.long)
    atoms = torch.randint(4, 30, (B, n_atoms), generator=g)
    xyz = torch.randn(B, n_atoms, 3, generator=g) * 2.0
    idx = torch.arange(n_atoms)[None, :]
    amask = idx < n[:, None]
    xyz = xyz * amask[..., None]
    xyz = xyz - (xyz.sum(1, keepdim=True) / n[:, None, None].float()) * amask[..., None]
    tokens = torch.zeros(B, L, dtype=torch.long)
    tokens[:, 0] = 1
    tokens[:, 1:n_atoms + 1] = atoms * amask
    tokens[torch.arange(B), n + 1] = 2
    coord = torch.zeros(B, L, 3)
    coord[:, 1:n_atoms + 1] = xyz
    valid = tokens.ne(0)
    dist = torch.cdist(coord, coord) * (valid[:, :, None] & valid[:, None, :])
    et = (tokens[:, :, None] * n_dict + tokens[:, None, :]) * (valid[:, :, None] & valid[:, None, :])
    return tokens, dist.float().contiguous(), et.contiguous(), coord


def featurise(src_tokens, src_coord, n_dict=31, pad_idx=0):
    """src_distance (B,L,L) f32 and src_edge_type (B,L,L) int64 computed ON THE DEVICE from the tokens (B,L) and the
    centred coordinates (B,L,3) of a padded batch -- bit-exact with data/conformer.py:205-218 + the zero padding of
    utils/util.py:41-105, so a step uploads B*L*20 bytes instead of B*L*L*12.  n_dict is
    len(dictionary) (31 with [MASK], models/mm_model.py:435-437)."""
    _lib.require_cuda(src_tokens)
    tok = src_tokens.long().contiguous()
    xyz = src_coord.float().contiguous()
    B, L = tok.shape
    if xyz.shape != (B, L, 3):
        raise ValueError("featurise: src_coord must be (B, L, 3) matching src_tokens (B, L)")
    dist = torch.empty((B, L, L), device=tok.device, dtype=torch.float32)
    et = torch.empty((B, L, L), device=tok.device, dtype=torch.int64)
    call("mmdti_featurise", xyz, tok, i32(B), i32(L), i32(n_dict), i64(pad_idx), dist, et, stream_ptr())
    return dist, et
