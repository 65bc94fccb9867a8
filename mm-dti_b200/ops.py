"""torch.autograd.Function wrappers around the C ABI (include/mmdti_b200.h).

Every op requires CUDA tensors and the compiled library; there is no fallback."""
import threading

import os

import torch

from . import _lib, config, ops_gemm
from ._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64

_seed_lock = threading.Lock()
_seed_counter = 0
_seed_rank = 0


def set_seed_rank(rank):
    """Data parallel: fold the rank into every dropout seed, so that replicas which share torch's seed (identical
    weights) still draw independent masks for their shards of the global batch, like DDP with per-rank RNG state."""
    global _seed_rank
    _seed_rank = int(rank)


def next_seed():
    """Deterministic per-call dropout seed derived from torch's seed, the data-parallel rank (set_seed_rank) and a
    call counter."""
    global _seed_counter
    with _seed_lock:
        _seed_counter += 1
        c = _seed_counter
    return (torch.initial_seed() * 0x9E3779B97F4A7C15 + c * 0xBF58476D1CE4E5B9
            + _seed_rank * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF


def reset_seed_counter():
    global _seed_counter
    _seed_counter = 0


def pair_ld(L):
    """Row stride Lp of the padded (B,H,L,Lp) pair layout (include/mmdti_b200.h, PAIR LAYOUT)."""
    fn = _lib.lib().mmdti_pair_ld
    lp = fn(int(L))
    if lp <= 0:
        raise _lib.MMDTIError("sequence length L=%d exceeds the supported maximum 264" % L)
    return lp


class PairPadFn(torch.autograd.Function):
    """dense (B*H, L, L) any float dtype -> padded (B, H, L, Lp) pair dtype (-inf padding)."""

    @staticmethod
    def forward(ctx, dense, B, H, L, out_dtype):
        _lib.require_cuda(dense)
        dense = dense.contiguous()
        out = torch.empty((B, H, L, pair_ld(L)), device=dense.device, dtype=out_dtype)
        call("mmdti_pair_pad", dense, out, i32(B * H), i32(L), i32(DTYPE_CODE[dense.dtype]), i32(DTYPE_CODE[out_dtype]),
             stream_ptr())
        ctx.meta = (B, H, L, dense.dtype, dense.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, L, in_dtype, shape = ctx.meta
        g = g.contiguous()
        out = torch.empty(shape, device=g.device, dtype=torch.float32)
        call("mmdti_pair_unpad", g, out, i32(B * H), i32(L), i32(DTYPE_CODE[g.dtype]), i32(_lib.F32), stream_ptr())
        return out.to(in_dtype), None, None, None, None


def pair_unpad(padded, L, out_dtype=torch.float32):
    """padded (B,H,L,Lp) -> dense (B*H, L, L) (no autograd; for outputs / tests)."""
    B, H = padded.shape[0], padded.shape[1]
    out = torch.empty((B * H, L, L), device=padded.device, dtype=out_dtype)
    call("mmdti_pair_unpad", padded.contiguous(), out, i32(B * H), i32(L), i32(DTYPE_CODE[padded.dtype]),
         i32(DTYPE_CODE[out_dtype]), stream_ptr())
    return out


# =========================================================================== K1
class PairBiasFn(torch.autograd.Function):
    """dist (B,L,L) f32, edge_type (B,L,L) int64 -> pair bias in the padded (B,H,L,Lp) layout.
    Reference: models/mm_model.py:553-556 (gbf -> gbf_proj -> permute -> contiguous)."""

    @staticmethod
    def forward(ctx, dist, edge_type, means, stds, mul, bias, w1, b1, w2, b2, key_pad, out_dtype, fp32_math):
        _lib.require_cuda(dist, edge_type, means, w1)
        B, L = dist.shape[0], dist.shape[1]
        K, H, E = means.numel(), w2.shape[0], mul.numel()
        dist = dist.contiguous().float()
        edge_type = edge_type.contiguous().long()
        ps = [t.detach().contiguous().float() for t in (means, stds, mul, bias, w1, b1, w2, b2)]
        kp = key_pad.contiguous().to(torch.uint8) if key_pad is not None else None
        out = torch.empty((B, H, L, pair_ld(L)), device=dist.device, dtype=out_dtype)
        call("mmdti_pair_bias_fwd", dist, edge_type, *ps, kp, out, i32(B), i32(L), i32(K), i32(H), i32(E),
             i32(DTYPE_CODE[out_dtype]), i32(1 if fp32_math else 0), stream_ptr())
        ctx.save_for_backward(dist, edge_type, *ps)
        ctx.fp32_math = fp32_math
        ctx.dims = (B, L, K, H, E)
        return out

    @staticmethod
    def backward(ctx, d_out):
        dist, edge_type, means, stds, mul, bias, w1, b1, w2, b2 = ctx.saved_tensors
        B, L, K, H, E = ctx.dims
        cdt = torch.float32 if ctx.fp32_math else torch.bfloat16
        dev = dist.device
        d_means = torch.zeros(K, device=dev)
        d_stds = torch.zeros(K, device=dev)
        d_mul = torch.zeros(E, device=dev)
        d_bias = torch.zeros(E, device=dev)
        d_w1 = torch.zeros(K, K, device=dev)
        d_b1 = torch.zeros(K, device=dev)
        d_w2 = torch.zeros(H, K, device=dev)
        d_b2 = torch.zeros(H, device=dev)
        d_out = d_out.contiguous()
        if not ctx.fp32_math:
            # production path: one fused tensor-core kernel, nothing of size pairs x 128 reaches HBM
            call("mmdti_pair_bias_bwd", d_out, dist, edge_type, means, stds, mul, bias, w1, b1, w2, d_means, d_stds, d_mul,
                 d_bias, d_w1, d_b1, d_w2, d_b2, i32(B), i32(L), i32(K), i32(H), i32(E), i32(DTYPE_CODE[d_out.dtype]),
                 stream_ptr())
            return (None, None, d_means.view(1, K), d_stds.view(1, K), d_mul.view(E, 1), d_bias.view(E, 1),
                    d_w1, d_b1, d_w2, d_b2, None, None, None)
        w1c, w2c, b1c = w1.to(cdt), w2.to(cdt), b1.to(cdt)
        # molecules per chunk: bound the (pairs,128) temporaries to ~256 MB
        per_mol = L * L * K * (4 if ctx.fp32_math else 2) * 6
        chunk = max(1, min(B, (256 << 20) // max(per_mol, 1)))
        for b0 in range(0, B, chunk):
            b1_ = min(B, b0 + chunk)
            nb = b1_ - b0
            npairs = nb * L * L
            dist_c, et_c = dist[b0:b1_], edge_type[b0:b1_]
            G = torch.empty((npairs, K), device=dev, dtype=cdt)
            call("mmdti_gauss_basis", dist_c, et_c, means, stds, mul, bias, G, i64(npairs), i32(K), i32(E),
                 i32(DTYPE_CODE[cdt]), stream_ptr())
            dO = torch.empty((npairs, H), device=dev, dtype=cdt)
            call("mmdti_pair_to_rows", d_out[b0:b1_], dO, i32(nb), i32(H), i32(L), i32(DTYPE_CODE[d_out.dtype]),
                 i32(DTYPE_CODE[cdt]), stream_ptr())
            Z = torch.addmm(b1c, G, w1c.t())
            Hm = torch.nn.functional.gelu(Z)
            d_w2 += (dO.t() @ Hm).float()
            d_b2 += dO.float().sum(0)
            dH = dO @ w2c
            dZ = torch.ops.aten.gelu_backward(dH, Z)
            d_w1 += (dZ.t() @ G).float()
            d_b1 += dZ.float().sum(0)
            dG = (dZ @ w1c).contiguous()
            call("mmdti_gauss_param_grad", dG, dist_c, et_c, means, stds, mul, bias, d_means, d_stds, d_mul,
                 d_bias, i64(npairs), i32(K), i32(E), i32(DTYPE_CODE[cdt]), stream_ptr())
        return (None, None, d_means.view(1, K), d_stds.view(1, K), d_mul.view(E, 1), d_bias.view(E, 1),
                d_w1, d_b1, d_w2, d_b2, None, None, None)


def pair_bias(dist, edge_type, means, stds, mul, bias, w1, b1, w2, b2, key_pad=None):
    return PairBiasFn.apply(dist, edge_type, means, stds, mul, bias, w1, b1, w2, b2, key_pad,
                            config.pair_dtype(), config.fp32_mode())


class GaussBasisFn(torch.autograd.Function):
    """Stand-alone GaussianLayer.forward (models/mm_model.py:254-269) -> (B,L,L,K) f32."""

    @staticmethod
    def forward(ctx, dist, edge_type, means, stds, mul, bias):
        _lib.require_cuda(dist, edge_type, means)
        K, E = means.numel(), mul.numel()
        dist = dist.contiguous().float()
        edge_type = edge_type.contiguous().long()
        ps = [t.detach().contiguous().float() for t in (means, stds, mul, bias)]
        out = torch.empty(tuple(dist.shape) + (K,), device=dist.device, dtype=torch.float32)
        call("mmdti_gauss_basis", dist, edge_type, *ps, out, i64(dist.numel()), i32(K), i32(E), i32(_lib.F32),
             stream_ptr())
        ctx.save_for_backward(dist, edge_type, *ps)
        return out

    @staticmethod
    def backward(ctx, dG):
        dist, edge_type, means, stds, mul, bias = ctx.saved_tensors
        K, E = means.numel(), mul.numel()
        dev = dist.device
        d_means, d_stds = torch.zeros(K, device=dev), torch.zeros(K, device=dev)
        d_mul, d_bias = torch.zeros(E, device=dev), torch.zeros(E, device=dev)
        dG = dG.contiguous().float()
        call("mmdti_gauss_param_grad", dG, dist, edge_type, means, stds, mul, bias, d_means, d_stds, d_mul, d_bias,
             i64(dist.numel()), i32(K), i32(E), i32(_lib.F32), stream_ptr())
        return None, None, d_means.view(1, K), d_stds.view(1, K), d_mul.view(E, 1), d_bias.view(E, 1)


class PairMaskFillFn(torch.autograd.Function):
    """In-place fill of padded KEY columns as an autograd node: ``mark_dirty`` bumps the tensor's version counter the
    way the reference's ``masked_fill_`` does (an upstream op that saved the tensor then raises instead of silently
    using mutated values), and the gradient is zeroed at the filled positions."""

    @staticmethod
    def forward(ctx, pair, kp, B, H, L, ld, fill):
        call("mmdti_pair_mask_fill", pair, kp, i32(B), i32(H), i32(L), i32(ld), i32(DTYPE_CODE[pair.dtype]), f32(fill),
             stream_ptr())
        ctx.mark_dirty(pair)
        ctx.save_for_backward(kp)
        ctx.dims = (B, H, L, ld)
        return pair

    @staticmethod
    def backward(ctx, g):
        (kp,) = ctx.saved_tensors
        B, H, L, ld = ctx.dims
        m = kp.bool()
        if ld != L:
            m = torch.nn.functional.pad(m, (0, ld - L), value=False)
        return g.reshape(B, H, L, ld).masked_fill(m[:, None, None, :], 0).reshape(g.shape), None, None, None, None, None, None


def pair_mask_fill_(pair, key_pad, fill=float("-inf")):
    """In place: pair[b,h,:,j] = fill where key_pad[b,j] (models/transformers.py:122-132).
    ``pair`` is either the reference's dense (B*H,L,L) tensor or the padded (B,H,L,Lp) layout.  Returns ``pair``
    (the same tensor object; use the return value so that autograd sees the in-place node)."""
    _lib.require_cuda(pair, key_pad)
    B, L = key_pad.shape
    if not pair.is_contiguous():
        raise _lib.MMDTIError("pair_mask_fill_: the pair tensor must be contiguous")
    ld = pair.shape[-1]
    if ld != L and ld != pair_ld(L):
        raise _lib.MMDTIError("pair_mask_fill_: last dim %d is neither L=%d nor Lp=%d" % (ld, L, pair_ld(L)))
    H = pair.numel() // (B * L * ld)
    kp = key_pad.contiguous().to(torch.uint8)
    return PairMaskFillFn.apply(pair, kp, B, H, L, ld, float(fill))


# =========================================================================== K2
class PairAttnFn(torch.autograd.Function):
    """qkv (B*L, 3*H*8) [q|k|v], pair_in (B,H,L,Lp) -> o (B*L, H*8), pair_out (B,H,L,Lp).
    Reference: Uni-Core SelfMultiheadAttention(return_attn=True) core via
    models/transformers.py:136-139."""

    @staticmethod
    def forward(ctx, qkv, pair_in, B, H, L, scale, dropout_p, seed, inplace_pair):
        _lib.require_cuda(qkv, pair_in)
        D = H * 8
        if qkv.shape != (B * L, 3 * D) or not qkv.is_contiguous():
            raise _lib.MMDTIError("pair_attn: qkv must be a contiguous (B*L, 3*H*8) tensor")
        if pair_in.numel() != B * H * L * pair_ld(L) or not pair_in.is_contiguous():
            raise _lib.MMDTIError("pair_attn: pair must be a contiguous padded (B,H,L,Lp=%d) tensor" % pair_ld(L))
        o = torch.empty((B * L, D), device=qkv.device, dtype=qkv.dtype)
        pair_out = pair_in if inplace_pair else torch.empty_like(pair_in)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        call("mmdti_pair_attn_fwd", q, k, v, i64(3 * D), pair_in, pair_out, o, i64(D), i32(B), i32(H), i32(L),
             f32(scale), f32(dropout_p), u64(seed), i32(DTYPE_CODE[qkv.dtype]), i32(DTYPE_CODE[pair_in.dtype]),
             stream_ptr())
        if inplace_pair:
            ctx.mark_dirty(pair_in)
        ctx.save_for_backward(qkv, pair_out, o)
        ctx.cfg = (B, H, L, scale, dropout_p, seed)
        ctx.set_materialize_grads(False)
        return o, pair_out

    @staticmethod
    def backward(ctx, d_o, d_pair_out):
        qkv, s, o = ctx.saved_tensors
        B, H, L, scale, dropout_p, seed = ctx.cfg
        D = H * 8
        gdt = s.dtype          # the pair gradient is carried in the pair tensor's own dtype
        if d_o is None:
            d_o = torch.zeros_like(o)
        d_o = d_o.contiguous()
        if d_o.dtype != o.dtype:
            d_o = d_o.to(o.dtype)
        if d_pair_out is not None:
            d_pair_out = d_pair_out.contiguous()
            if d_pair_out.dtype != gdt:
                d_pair_out = d_pair_out.to(gdt)
        d_pair_in = torch.empty(s.shape, device=s.device, dtype=gdt)
        dqkv = torch.empty_like(qkv)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
        call("mmdti_pair_attn_bwd", q, k, v, i64(3 * D), s, o, d_o, i64(D), d_pair_out, d_pair_in, dq, dk, dv,
             i64(3 * D), i32(B), i32(H), i32(L), f32(scale), f32(dropout_p), u64(seed), i32(DTYPE_CODE[qkv.dtype]),
             i32(DTYPE_CODE[s.dtype]), i32(DTYPE_CODE[gdt]), stream_ptr())
        return dqkv, d_pair_in, None, None, None, None, None, None, None


def pair_attention(qkv, pair_in, B, H, L, scale, dropout_p=0.0, seed=0, inplace_pair=False):
    return PairAttnFn.apply(qkv, pair_in, B, H, L, scale, dropout_p, seed, inplace_pair)


def attn_dropout_mask(B, H, L, dropout_p, seed, device="cuda"):
    """Debug/test export of the keep mask used by pair_attention for `seed`."""
    keep = torch.empty((B, H, L, L), device=device, dtype=torch.uint8)
    call("mmdti_pair_attn_dropout_mask", keep, i32(B), i32(H), i32(L), f32(dropout_p), u64(seed), stream_ptr())
    return keep.bool()


class PairOutputsFn(torch.autograd.Function):
    """(pair_first, pair_last) padded (B,H,L,Lp) -> pair (B,L,L,H) f32, delta (B,L,L,H) f32
    (models/transformers.py:163-172)."""

    @staticmethod
    def forward(ctx, pair_first, pair_last, B, H, L):
        pair = torch.empty((B, L, L, H), device=pair_last.device, dtype=torch.float32)
        delta = torch.empty_like(pair)
        call("mmdti_pair_outputs", pair_first, pair_last, pair, delta, i32(B), i32(H), i32(L),
             i32(DTYPE_CODE[pair_last.dtype]), stream_ptr())
        ctx.save_for_backward(pair_last)
        ctx.dims = (B, H, L, pair_first.dtype)
        ctx.set_materialize_grads(False)
        return pair, delta

    @staticmethod
    def backward(ctx, d_pair, d_delta):
        (pair_last,) = ctx.saved_tensors
        B, H, L, first_dtype = ctx.dims
        g_last = None
        if d_pair is not None:
            g_last = d_pair
        if d_delta is not None:
            g_last = d_delta if g_last is None else g_last + d_delta
        if g_last is None:
            return None, None, None, None, None
        Lp = pair_last.shape[-1]
        valid = torch.isfinite(pair_last[..., :L])
        pad = (0, Lp - L)
        g_last = torch.nn.functional.pad((g_last.permute(0, 3, 1, 2) * valid).to(pair_last.dtype), pad)
        g_first = None
        if d_delta is not None:
            g_first = torch.nn.functional.pad((-(d_delta.permute(0, 3, 1, 2) * valid)).to(first_dtype), pad).contiguous()
        return g_first, g_last.contiguous(), None, None, None


# =========================================================================== fused encoder layer
def _mm_f32(a, b):
    """a @ b with an fp32 result (weight gradients keep full accumulator precision)."""
    if a.dtype == torch.float32:
        return torch.mm(a, b)
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:
        return torch.mm(a, b).float()


# Which dense projections of the encoder layer run on the own tcgen05 GEMMs with fused epilogues (csrc/gemm_tc.cu)
# instead of library GEMMs + separate elementwise kernels.  MMDTI_FUSED=all | none | comma list of:
#   in    in_proj forward (+bias)                         out   out_proj forward + dropout + residual + LayerNorm-2
#   fc1   fc1 forward + bias + GELU                       fc2   fc2 forward + dropout + residual (+ next LayerNorm-1)
#   dfc2  fc2 dgrad + GELU backward + db_fc1              dfc1  fc1 dgrad + LayerNorm-2 backward + dropout backward
#   dout  out_proj dgrad                                  din   in_proj dgrad
#   wgrad the four weight gradients
FUSED_ALL = ("in", "out", "fc1", "fc2", "dfc2", "dfc1", "dout", "din", "wgrad")


def _parse_fused(spec):
    spec = spec.strip().lower()
    if spec in ("all", "1", ""):
        return frozenset(FUSED_ALL)
    if spec in ("none", "0"):
        return frozenset()
    names = frozenset(x.strip() for x in spec.split(",") if x.strip())
    bad = names - frozenset(FUSED_ALL)
    if bad:
        raise ValueError("MMDTI_FUSED: unknown entries %s (known: %s)" % (sorted(bad), ", ".join(FUSED_ALL)))
    return names


fused_gemms = _parse_fused(os.environ.get("MMDTI_FUSED", "all"))

_side_streams = {}
overlap_wgrad = True      # run weight-gradient GEMMs / bias column sums of a layer on a side stream (off the critical path)

def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def layernorm_fwd(x2d, w, b, out_dtype, eps=1e-5):
    rows, D = x2d.shape
    y = torch.empty((rows, D), device=x2d.device, dtype=out_dtype)
    stats = torch.empty((2, rows), device=x2d.device, dtype=torch.float32)
    call("mmdti_layernorm_fwd", x2d, w, b, y, stats[0], stats[1], i32(rows), i32(D), f32(eps), i32(DTYPE_CODE[out_dtype]),
         stream_ptr())
    return y, stats


class LayerNormFn(torch.autograd.Function):
    """Stand-alone LayerNorm (unicore.modules.LayerNorm, eps 1e-5) on the mmdti kernels: fp32 in / fp32 out.
    Used for the encoder's emb_layer_norm / final_layer_norm (models/transformers.py:69-78,114,160)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        _lib.require_cuda(x)
        D = x.shape[-1]
        x2d = x.detach().reshape(-1, D).contiguous().float()
        wd, bd = w.detach().float().contiguous(), b.detach().float().contiguous()
        y = torch.empty_like(x2d)
        st = torch.empty((2, x2d.shape[0]), device=x.device, dtype=torch.float32)
        call("mmdti_layernorm_fwd", x2d, wd, bd, y, st[0], st[1], i32(x2d.shape[0]), i32(D), f32(eps), i32(_lib.F32), stream_ptr())
        ctx.save_for_backward(x2d, wd, st)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2d, w, st = ctx.saved_tensors
        rows, D = x2d.shape
        dy2 = dy.reshape(rows, D).contiguous().float()
        dx = torch.empty_like(x2d)
        dwb = torch.zeros(2 * D, device=x2d.device, dtype=torch.float32)
        call("mmdti_layernorm_bwd", dy2, x2d, w, st[0], st[1], None, dx, dwb[:D], dwb[D:], i32(rows), i32(D), i32(_lib.F32),
             stream_ptr())
        return dx.view(ctx.shape), dwb[:D], dwb[D:], None


class TokenEmbeddingFn(torch.autograd.Function):
    """nn.Embedding lookup whose backward is one index_add (atomics) instead of torch's sort-based path:
    the Uni-Mol dictionary has 31 rows, so the sorted segmented reduction is pure overhead."""

    @staticmethod
    def forward(ctx, tokens, weight, padding_idx):
        ctx.save_for_backward(tokens)
        ctx.meta = (weight.shape, padding_idx)
        return weight[tokens]

    @staticmethod
    def backward(ctx, g):
        (tokens,) = ctx.saved_tensors
        shape, padding_idx = ctx.meta
        gw = torch.zeros(shape, device=g.device, dtype=g.dtype)
        gw.index_add_(0, tokens.reshape(-1), g.reshape(-1, shape[1]))
        if padding_idx is not None:
            gw[padding_idx].zero_()
        return None, gw, None


class EncoderLayerFn(torch.autograd.Function):
    """One pre-LN Uni-Core TransformerEncoderLayer with return_attn=True (SURVEY.md Appendix A;
    call site models/transformers.py:136-139) as a single autograd node with a hand-written
    backward.  bf16 mode: the four dense projections, forward / dgrad / wgrad, are the tcgen05 GEMMs of csrc/gemm_tc.cu whose
    epilogues carry LayerNorm, dropout + residual, GELU and the bias / LayerNorm column sums (ops_gemm; `MMDTI_FUSED` selects
    a subset); fp32 validation mode: fp32 library GEMMs around the stand-alone elementwise kernels.  K2 is the pair-biased
    attention.
    x (B,L,D) f32 residual stream; pair_in padded (B,H,L,Lp)."""

    @staticmethod
    def forward(ctx, x, pair_in, ln1_w, ln1_b, w_in, b_in, w_out, b_out, ln2_w, ln2_b, w_fc1, b_fc1, w_fc2, b_fc2,
                lowp, cfg, h1_in=None, st1_in=None, nxt_w=None, nxt_b=None, link_in=None, link_out=None):
        """Cross-layer chaining (optional): ``h1_in, st1_in`` = this layer's LayerNorm-1 output and statistics, already
        produced by the PREVIOUS layer; ``nxt_w, nxt_b`` = the NEXT layer's LayerNorm-1 parameters, in which case the
        final dropout+residual is fused with that LayerNorm (one kernel instead of two, forward and backward) and the
        function returns (x2, pair_out, h_next, st_next).  ``link_in`` / ``link_out``: plain dicts shared with the previous /
        next layer of the chain; the forward leaves the dropout stream of this layer's final dropout in ``link_out`` so
        that the NEXT layer's backward can fuse its in_proj dgrad + LayerNorm-1 backward with that dropout's backward
        (one GEMM epilogue, csrc/gemm_tc.cu EPI_LNBWD_DROP) and hand the results back through the same dict."""
        B, H, L, scale, p_attn, p_drop, seeds, dt = cfg
        _lib.require_cuda(x, pair_in)
        D = x.shape[-1]
        rows = B * L
        code = DTYPE_CODE[dt]
        sp = stream_ptr()
        if lowp is None:
            lowp = [t.detach().to(dt) for t in (w_in, b_in, w_out, b_out, w_fc1, b_fc1, w_fc2, b_fc2)]
        w_in_l, b_in_l, w_out_l, b_out_l, w_fc1_l, b_fc1_l, w_fc2_l, b_fc2_l = lowp
        x2d = x.detach().reshape(rows, D).contiguous().float()
        ln1_wd, ln1_bd, ln2_wd, ln2_bd = (t.detach().float().contiguous() for t in (ln1_w, ln1_b, ln2_w, ln2_b))
        pair_in = pair_in.detach()
        chain_in = h1_in is not None
        if chain_in:
            h1, st1 = h1_in.detach(), st1_in.detach()
        else:
            h1, st1 = layernorm_fwd(x2d, ln1_wd, ln1_bd, dt)
        # own tcgen05 GEMMs with fused epilogues where the shapes allow it (bf16, feature dims multiples of 64, D <= 512)
        fz = fused_gemms if (dt == torch.bfloat16 and ops_gemm.supported(D, w_fc1_l.shape[0])) else frozenset()
        qkv = ops_gemm.gemm_bias(h1, w_in_l, b_in_l) if "in" in fz else torch.addmm(b_in_l, h1, w_in_l.t())
        o = torch.empty((rows, D), device=x.device, dtype=dt)
        pair_out = torch.empty_like(pair_in)
        call("mmdti_pair_attn_fwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pair_in, pair_out, o, i64(D),
             i32(B), i32(H), i32(L), f32(scale), f32(p_attn), u64(seeds[0]), i32(code), i32(DTYPE_CODE[pair_in.dtype]), sp)
        if "out" in fz:
            # out_proj + bias + dropout + residual + LayerNorm-2 in one kernel (same dropout stream as the unfused kernel)
            x1, h2, st2 = ops_gemm.gemm_dropres_ln(o, w_out_l, b_out_l, x2d, ln2_wd, ln2_bd, p_drop, seeds[1])
        else:
            a = torch.addmm(b_out_l, o, w_out_l.t())
            x1 = torch.empty_like(x2d)
            h2 = torch.empty((rows, D), device=x.device, dtype=dt)
            st2 = torch.empty((2, rows), device=x.device, dtype=torch.float32)
            call("mmdti_dropres_layernorm_fwd", x2d, a, x1, ln2_wd, ln2_bd, h2, st2[0], st2[1], i32(rows), i32(D), f32(1e-5),
                 f32(p_drop), u64(seeds[1]), i32(code), i32(code), sp)
        # with both fc1 and its backward on the fused GEMMs the forward stores gelu'(z) in place of z: the backward
        # epilogue then is one multiply per element instead of re-evaluating erf / exp on every element
        z_is_grad = "fc1" in fz and "dfc2" in fz
        if "fc1" in fz:
            z, u = ops_gemm.gemm_bias_gelu(h2, w_fc1_l, b_fc1_l, store_grad=z_is_grad)
        else:
            z = torch.addmm(b_fc1_l, h2, w_fc1_l.t())
            u = torch.empty_like(z)
            call("mmdti_gelu_fwd", z, u, i64(z.numel()), i32(code), sp)
        chain_out = nxt_w is not None
        nxt_wd = h_next = st_next = None
        if chain_out:
            nxt_wd, nxt_bd = nxt_w.detach().float().contiguous(), nxt_b.detach().float().contiguous()
        if "fc2" in fz:
            x2, h_next, st_next = ops_gemm.gemm_dropres_ln(u, w_fc2_l, b_fc2_l, x1, nxt_wd, nxt_bd if chain_out else None,
                                                           p_drop, seeds[2])
        else:
            f = torch.addmm(b_fc2_l, u, w_fc2_l.t())
            x2 = torch.empty_like(x2d)
            if chain_out:
                h_next = torch.empty((rows, D), device=x.device, dtype=dt)
                st_next = torch.empty((2, rows), device=x.device, dtype=torch.float32)
                call("mmdti_dropres_layernorm_fwd", x1, f, x2, nxt_wd, nxt_bd, h_next, st_next[0], st_next[1], i32(rows), i32(D),
                     f32(1e-5), f32(p_drop), u64(seeds[2]), i32(code), i32(code), sp)
            else:
                call("mmdti_dropout_residual_fwd", x1, f, x2, i64(rows * D), f32(p_drop), u64(seeds[2]), i32(code), sp)
        ctx.save_for_backward(x2d, pair_out, st1, h1, qkv, o, x1, st2, h2, z, u, ln1_wd, ln2_wd, w_in_l, w_out_l,
                              w_fc1_l, w_fc2_l, *((x2, st_next, nxt_wd) if chain_out else ()))
        ctx.cfg = cfg
        ctx.chain = (chain_in, chain_out)
        ctx.fz = fz
        ctx.z_is_grad = z_is_grad
        ctx.links = (link_in if chain_in else None, link_out if chain_out else None)
        if chain_out and link_out is not None:
            link_out["p"], link_out["seed"] = p_drop, seeds[2]
        ctx.set_materialize_grads(False)
        if chain_out:
            ctx.mark_non_differentiable(st_next)
            return x2.view(B, L, D), pair_out, h_next, st_next
        return x2.view(B, L, D), pair_out

    @staticmethod
    def backward(ctx, dx2, dpair_out, dh_next=None, _dst_next=None):
        saved = ctx.saved_tensors
        (x2d, pair_out, st1, h1, qkv, o, x1, st2, h2, z, u, ln1_w, ln2_w, w_in_l, w_out_l, w_fc1_l, w_fc2_l) = saved[:17]
        chain_in, chain_out = ctx.chain
        B, H, L, scale, p_attn, p_drop, seeds, dt = ctx.cfg
        rows, D = x2d.shape
        F_ = z.shape[1]
        code = DTYPE_CODE[dt]
        sp = stream_ptr()
        dev = x2d.device
        if dx2 is None:
            dx2 = torch.zeros((rows, D), device=dev, dtype=torch.float32)
        dx2 = dx2.reshape(rows, D).contiguous().float()
        if dpair_out is not None:
            dpair_out = dpair_out.contiguous()
            if dpair_out.dtype != pair_out.dtype:
                dpair_out = dpair_out.to(pair_out.dtype)
        # one zeroed buffer for all reduced gradients of this layer
        red = torch.zeros(4 * D + 2 * D + F_ + 3 * D + 2 * D, device=dev, dtype=torch.float32)
        dw_ln1, db_ln1, dw_ln2, db_ln2 = red[0:D], red[D:2 * D], red[2 * D:3 * D], red[3 * D:4 * D]
        db_out, db_fc2 = red[4 * D:5 * D], red[5 * D:6 * D]
        db_fc1 = red[6 * D:6 * D + F_]
        db_in = red[6 * D + F_:9 * D + F_]
        dw_nxt, db_nxt = red[9 * D + F_:10 * D + F_], red[10 * D + F_:]
        # Weight gradients (and the in_proj bias column sums) are not on the critical path of the backward
        # chain: they go to a side stream and overlap the latency-bound kernels of the main stream (in a CUDA
        # graph they become parallel branches).  The layer joins the side stream before it returns.
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if overlap_wgrad else None

        def off_path(fn):
            if side is None:
                return fn()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = fn()
            out.record_stream(main)
            return out

        # ---- feed-forward block
        df = torch.empty((rows, D), device=dev, dtype=dt)
        link_in, link_out = ctx.links
        if chain_out and dh_next is not None and link_out is not None and link_out.pop("fused", False):
            # the next layer's backward already ran in_proj dgrad + LayerNorm-1 backward + THIS layer's final dropout
            # backward in one GEMM epilogue: dx2 is the total gradient at x2, dh_next carries df, the dict db_fc2
            df = dh_next.contiguous()
            db_fc2 = link_out.pop("db_fc2")
            g_nxt = (None, None)
        elif chain_out and dh_next is not None:
            # the NEXT layer's LayerNorm-1 backward fused with this layer's final dropout backward:
            # dxt = dx2 + dLN(dh_next) is the total gradient at x2, df = dropout'(dxt)
            x2, st_next, nxt_w = saved[17:20]
            dxt = torch.empty_like(x2d)
            call("mmdti_layernorm_bwd_dropout", dh_next.contiguous(), x2, nxt_w, st_next[0], st_next[1], dx2, dxt, dw_nxt, db_nxt, df,
                 db_fc2, i32(rows), i32(D), f32(p_drop), u64(seeds[2]), i32(code), sp)
            dx2 = dxt
            g_nxt = (dw_nxt, db_nxt)
        else:
            call("mmdti_dropout_bwd", dx2, df, db_fc2, i32(rows), i32(D), f32(p_drop), u64(seeds[2]), i32(code), sp)
            g_nxt = (None, None)
        fz = ctx.fz
        wgrad = (lambda dy, xin: ops_gemm.gemm_wgrad(dy, xin)) if "wgrad" in fz else (lambda dy, xin: _mm_f32(dy.t(), xin))
        dW_fc2 = off_path(lambda: wgrad(df, u))
        if "dfc2" in fz:
            dz = ops_gemm.gemm_dgrad_gelu(df, w_fc2_l, z, db_fc1, z_is_grad=ctx.z_is_grad)      # fc2 dgrad + GELU' + db_fc1 column sums
        else:
            du = torch.mm(df, w_fc2_l)
            dz = torch.empty_like(z)
            call("mmdti_gelu_bwd", du, z, dz, db_fc1, i32(rows), i32(F_), i32(code), sp)
        dW_fc1 = off_path(lambda: wgrad(dz, h2))
        # ---- attention block (LayerNorm-2 backward fused with the dropout backward of the attention output)
        if "dfc1" in fz:
            dx1, da = ops_gemm.gemm_dgrad_lnbwd(dz, w_fc1_l, x1, st2, ln2_w, dx2, dw_ln2, db_ln2, db_out, p_drop, seeds[1])
        else:
            dh2 = torch.mm(dz, w_fc1_l)
            dx1 = torch.empty_like(x2d)
            da = torch.empty((rows, D), device=dev, dtype=dt)
            call("mmdti_layernorm_bwd_dropout", dh2, x1, ln2_w, st2[0], st2[1], dx2, dx1, dw_ln2, db_ln2, da, db_out, i32(rows),
                 i32(D), f32(p_drop), u64(seeds[1]), i32(code), sp)
        dW_out = off_path(lambda: wgrad(da, o))
        d_o = ops_gemm.gemm_dgrad(da, w_out_l) if "dout" in fz else torch.mm(da, w_out_l)
        dqkv = torch.empty_like(qkv)
        dpair_in = torch.empty_like(pair_out)
        call("mmdti_pair_attn_bwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pair_out, o, d_o, i64(D),
             dpair_out, dpair_in, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], i64(3 * D), i32(B), i32(H), i32(L),
             f32(scale), f32(p_attn), u64(seeds[0]), i32(code), i32(DTYPE_CODE[pair_out.dtype]),
             i32(DTYPE_CODE[pair_out.dtype]), sp)
        def in_proj_grads():
            call("mmdti_colsum", dqkv, db_in, i32(rows), i32(3 * D), i32(code), stream_ptr())
            return wgrad(dqkv, h1)

        dW_in = off_path(in_proj_grads)
        if chain_in and "din" in fz and link_in is not None and "seed" in link_in:
            # in_proj dgrad + this layer's LayerNorm-1 backward + the PREVIOUS layer's final dropout backward in one GEMM
            # epilogue: dx = total gradient at the previous layer's x2, g_h1 = df of the previous layer, its db_fc2 in the dict
            db_fc2_prev = torch.zeros(D, device=dev, dtype=torch.float32)
            dx, g_h1 = ops_gemm.gemm_dgrad_lnbwd(dqkv, w_in_l, x2d, st1, ln1_w, dx1, dw_ln1, db_ln1, db_fc2_prev,
                                                 link_in["p"], link_in["seed"])
            link_in["db_fc2"], link_in["fused"] = db_fc2_prev, True
            g_ln1 = (dw_ln1, db_ln1)
        elif chain_in:
            dh1 = ops_gemm.gemm_dgrad(dqkv, w_in_l) if "din" in fz else torch.mm(dqkv, w_in_l)
            # LayerNorm-1 belongs to the previous layer's fused kernel: hand it dh1, return the residual gradient alone
            dx, g_ln1, g_h1 = dx1, (None, None), dh1
        else:
            dh1 = ops_gemm.gemm_dgrad(dqkv, w_in_l) if "din" in fz else torch.mm(dqkv, w_in_l)
            dx = dx1          # in place: dx = dx1 + dLN1
            call("mmdti_layernorm_bwd", dh1, x2d, ln1_w, st1[0], st1[1], dx1, dx, dw_ln1, db_ln1, i32(rows), i32(D), i32(code), sp)
            g_ln1, g_h1 = (dw_ln1, db_ln1), None
        if side is not None:
            main.wait_stream(side)
        return (dx.view(B, L, D), dpair_in, g_ln1[0], g_ln1[1], dW_in, db_in, dW_out, db_out, dw_ln2, db_ln2, dW_fc1, db_fc1,
                dW_fc2, db_fc2, None, None, g_h1, None, g_nxt[0], g_nxt[1], None, None)


def dropout_mask(n, p, seed, device="cuda"):
    """Debug/test export of the flat-tensor dropout keep mask for `seed`."""
    keep = torch.empty(n, device=device, dtype=torch.uint8)
    call("mmdti_dropout_mask", keep, i64(n), f32(p), u64(seed), stream_ptr())
    return keep.bool()
