"""Thin host wrappers of the tcgen05 projection GEMMs with fused epilogues (include/mmdti_b200.h, "dense projections
of the encoder layer"; csrc/gemm_tc.cu).  bf16 operands, fp32 accumulation; no fallback.  Reference call sites: Uni-Core
TransformerEncoderLayer via models/transformers.py:82-91,136-139."""
import torch

from . import _lib
from ._lib import call, f32, i32, i64, stream_ptr, u64


def supported(D, F_):
    """Feature dims the fused path covers: multiples of 64 (TMA boxes of 64 elements), model dim <= 512 (whole rows
    per CTA pair in the LayerNorm epilogues)."""
    return D % 64 == 0 and F_ % 64 == 0 and D <= 512


def _chk(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.bfloat16 or t.stride(-1) != 1):
            raise _lib.MMDTIError("gemm_tc: operands must be bf16 with unit inner stride")


def gemm_bias(x, w, bias, out=None):
    """x (M,K) @ w (N,K)^T + bias -> (M,N) bf16"""
    _chk(x, w, bias)
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=torch.bfloat16) if out is None else out
    call("mmdti_gemm_bias", x, i64(x.stride(0)), w, i64(w.stride(0)), bias, y, i64(y.stride(0)), i32(M), i32(N), i32(K), stream_ptr())
    return y


def gemm_bias_gelu(x, w, bias, store_grad=False):
    """-> (z, u): z = x w^T + bias, u = gelu(z); store_grad=True: the first result is gelu'(z) instead of z"""
    _chk(x, w, bias)
    M, K = x.shape
    N = w.shape[0]
    z = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    u = torch.empty_like(z)
    call("mmdti_gemm_bias_gelu", x, i64(x.stride(0)), w, i64(w.stride(0)), bias, z, i64(N), u, i64(N), i32(M), i32(N), i32(K),
         i32(1 if store_grad else 0), stream_ptr())
    return z, u


def gemm_dropres_ln(x, w, bias, res, ln_w, ln_b, p, seed, eps=1e-5):
    """xo = res + dropout(x w^T + bias); y = LayerNorm(xo) -> (xo f32, y bf16 | None, stats (2, M) | None)"""
    _chk(x, w, bias)
    M, K = x.shape
    N = w.shape[0]
    xo = torch.empty((M, N), device=x.device, dtype=torch.float32)
    if ln_w is not None:
        y = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
        st = torch.empty((2, M), device=x.device, dtype=torch.float32)
        call("mmdti_gemm_dropres_ln", x, i64(x.stride(0)), w, i64(w.stride(0)), bias, res, xo, ln_w, ln_b, y, st[0], st[1], i32(M), i32(N),
             i32(K), f32(eps), f32(p), u64(seed), stream_ptr())
        return xo, y, st
    call("mmdti_gemm_dropres_ln", x, i64(x.stride(0)), w, i64(w.stride(0)), bias, res, xo, None, None, None, None, None, i32(M), i32(N),
         i32(K), f32(eps), f32(p), u64(seed), stream_ptr())
    return xo, None, None


def gemm_dgrad(dy, w, out=None):
    """dy (M,N) @ w (N,K) -> (M,K) bf16"""
    _chk(dy, w)
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), device=dy.device, dtype=torch.bfloat16) if out is None else out
    call("mmdti_gemm_dgrad", dy, i64(dy.stride(0)), w, i64(w.stride(0)), dx, i64(dx.stride(0)), i32(M), i32(N), i32(K), stream_ptr())
    return dx


def gemm_dgrad_gelu(dy, w, z, dbias, z_is_grad=False):
    """dz = (dy @ w) * gelu'(z); dbias += colsum(dz).  z_is_grad=True: ``z`` already holds gelu'(z)"""
    _chk(dy, w, z)
    M, N = dy.shape
    K = w.shape[1]
    dz = torch.empty((M, K), device=dy.device, dtype=torch.bfloat16)
    call("mmdti_gemm_dgrad_gelu", dy, i64(dy.stride(0)), w, i64(w.stride(0)), z, i64(z.stride(0)), dz, i64(K), dbias, i32(M), i32(N), i32(K),
         i32(1 if z_is_grad else 0), stream_ptr())
    return dz


def gemm_dgrad_lnbwd(dy, w, x, stats, ln_w, dx_add, dw, db, dbias, p, seed):
    """dh = dy @ w; dx = dx_add + LN'(dh); da = dropout'(dx) -> (dx f32, da bf16); dw / db / dbias accumulated"""
    _chk(dy, w)
    M, N = dy.shape
    K = w.shape[1]
    dx = torch.empty((M, K), device=dy.device, dtype=torch.float32)
    da = torch.empty((M, K), device=dy.device, dtype=torch.bfloat16)
    call("mmdti_gemm_dgrad_lnbwd", dy, i64(dy.stride(0)), w, i64(w.stride(0)), x, stats[0], stats[1], ln_w, dx_add, dx, dw, db, da, dbias,
         i32(M), i32(N), i32(K), f32(p), u64(seed), stream_ptr())
    return dx, da


def gemm_wgrad(dy, x, out=None, accumulate=False):
    """dy (M,N)^T @ x (M,K) -> (N,K) f32"""
    _chk(dy, x)
    M, N = dy.shape
    K = x.shape[1]
    dw = torch.empty((N, K), device=dy.device, dtype=torch.float32) if out is None else out
    call("mmdti_gemm_wgrad", dy, i64(dy.stride(0)), x, i64(x.stride(0)), dw, i64(dw.stride(0)), i32(M), i32(N), i32(K),
         i32(1 if accumulate else 0), stream_ptr())
    return dw


class LinearGeluFn(torch.autograd.Function):
    """u = gelu(x W^T + b) on the tcgen05 GEMMs: x (M,K) any float dtype, W (N,K), b (N) fp32 parameters -> u (M,N) bf16.
    Forward = mmdti_gemm_bias_gelu; backward = GELU' (+ bias column sums) and the two gradient GEMMs.  Used by the InfoNCE
    projection heads (models/infonce.py:24-33: Linear -> GELU -> Linear over every token)."""

    @staticmethod
    def forward(ctx, x, w, b):
        _lib.require_cuda(x, w)
        xb = x.detach().to(torch.bfloat16).contiguous()
        wb = w.detach().to(torch.bfloat16).contiguous()
        bb = b.detach().to(torch.bfloat16).contiguous()
        z, u = gemm_bias_gelu(xb, wb, bb)
        ctx.save_for_backward(xb, wb, z)
        ctx.x_dtype = x.dtype
        ctx.needs = (x.requires_grad, w.requires_grad, b.requires_grad)
        return u

    @staticmethod
    def backward(ctx, du):
        xb, wb, z = ctx.saved_tensors
        M, N = z.shape
        du = du.contiguous().to(torch.bfloat16)
        dz = torch.empty_like(z)
        db = torch.zeros(N, device=z.device, dtype=torch.float32)
        call("mmdti_gelu_bwd", du, z, dz, db, i32(M), i32(N), i32(_lib.BF16), stream_ptr())
        dw = gemm_wgrad(dz, xb) if ctx.needs[1] else None
        dx = gemm_dgrad(dz, wb).to(ctx.x_dtype) if ctx.needs[0] else None
        return dx, dw, db if ctx.needs[2] else None


def linear_gelu(x, w, b):
    return LinearGeluFn.apply(x, w, b)
