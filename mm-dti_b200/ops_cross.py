"""Cross-modal fusion on the mmdti kernels: one post-LN ``BertCrossAttentionLayer``
(models/mm_module.py:549-560,607-620) as a single autograd node with a hand-written backward, the input dropout of
``CrossAttentionModel.forward`` (models/mm_model.py:386-391) and the masked mean pooling of ``MM_Model.forward``
(models/mm_model.py:571-576).

bf16 mode: the six projections run on the tcgen05 GEMMs of csrc/gemm_tc.cu with their epilogues (bias; dropout + residual +
LayerNorm; bias + GELU; GELU' + bias sums; LayerNorm' + dropout' + bias sums; split-K weight gradients), the attention on
csrc/cross_attn.cu.  fp32 mode (validation): library fp32 GEMMs around the same elementwise / attention kernels.  No other
fallback: a missing library raises."""
import math

import torch

from . import _lib, ops, ops_gemm
from ._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64


def cross_attn_dropout_mask(B, H, Lq, Lk, p, seed, device="cuda"):
    """Debug/test export of the keep mask of the attention-probability dropout for ``seed``: (B, H, Lq, Lk) bool."""
    keep = torch.empty((B, H, Lq, Lk), device=device, dtype=torch.uint8)
    call("mmdti_cross_attn_dropout_mask", keep, i32(B), i32(H), i32(Lq), i32(Lk), f32(p), u64(seed), stream_ptr())
    return keep.bool()


def _convert(src, dst_dtype, add=None, shape=None, copy=False):
    """dst = float(src) (+ add, fp32) in one vectorised pass (mmdti_convert_add); the no-op cases return src unless ``copy``.
    ``shape``: shape of the result (same element count) -- a fresh base tensor, not a view."""
    if add is None and src.dtype == dst_dtype and not copy:
        return src
    src = src.contiguous()
    dst = torch.empty(src.shape if shape is None else shape, device=src.device, dtype=dst_dtype)
    call("mmdti_convert_add", src, i32(DTYPE_CODE[src.dtype]), add, dst, i32(DTYPE_CODE[dst_dtype]), i64(src.numel()), stream_ptr())
    return dst


def _attn_fwd(q, kv, D, mask2, B, H, Lq, Lk, scale, p, seed):
    o = torch.empty((B * Lq, D), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, H, Lq), device=q.device, dtype=torch.float32)
    call("mmdti_cross_attn_fwd", q, i64(q.stride(0)), kv[:, :D], kv[:, D:], i64(kv.stride(0)), mask2, o, i64(D), lse, i32(B), i32(H),
         i32(Lq), i32(Lk), i32(D // H), f32(scale), f32(p), u64(seed), i32(DTYPE_CODE[q.dtype]), stream_ptr())
    return o, lse


def _attn_bwd(q, kv, D, mask2, o, d_o, lse, B, H, Lq, Lk, scale, p, seed):
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    delta = torch.empty_like(lse)
    call("mmdti_cross_attn_bwd", q, i64(q.stride(0)), kv[:, :D], kv[:, D:], i64(kv.stride(0)), mask2, o, d_o, i64(D), lse, delta, dq,
         i64(dq.stride(0)), dkv[:, :D], dkv[:, D:], i64(dkv.stride(0)), i32(B), i32(H), i32(Lq), i32(Lk), i32(D // H), f32(scale), f32(p),
         u64(seed), i32(DTYPE_CODE[q.dtype]), stream_ptr())
    return dq, dkv


class CrossAttnFn(torch.autograd.Function):
    """The attention core alone (tests): q (B*Lq, D), kv (B*Lk, 2D) -> context (B*Lq, D)."""

    @staticmethod
    def forward(ctx, q, kv, mask2, B, H, Lq, Lk, p, seed):
        _lib.require_cuda(q, kv)
        D = q.shape[1]
        scale = 1.0 / math.sqrt(D // H)
        q, kv = q.detach().contiguous(), kv.detach().contiguous()
        o, lse = _attn_fwd(q, kv, D, mask2, B, H, Lq, Lk, scale, p, seed)
        ctx.save_for_backward(q, kv, mask2, o, lse)
        ctx.cfg = (B, H, Lq, Lk, scale, p, seed)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, mask2, o, lse = ctx.saved_tensors
        B, H, Lq, Lk, scale, p, seed = ctx.cfg
        dq, dkv = _attn_bwd(q, kv, q.shape[1], mask2, o, d_o.contiguous().to(q.dtype), lse, B, H, Lq, Lk, scale, p, seed)
        return dq, dkv, None, None, None, None, None, None, None


class FlatDropoutFn(torch.autograd.Function):
    """nn.Dropout on the library's counter-based stream (mask of ops.dropout_mask(seed) over the flat tensor); fp32 in/out."""

    @staticmethod
    def forward(ctx, x, p, seed):
        _lib.require_cuda(x)
        x2 = x.detach().reshape(-1, x.shape[-1]).contiguous().float()
        y = torch.empty_like(x2)
        call("mmdti_dropout_bwd", x2, y, None, i32(x2.shape[0]), i32(x2.shape[1]), f32(p), u64(seed), i32(_lib.F32), stream_ptr())
        ctx.cfg = (p, seed)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.cfg
        d2 = dy.reshape(-1, dy.shape[-1]).contiguous().float()
        dx = torch.empty_like(d2)
        call("mmdti_dropout_bwd", d2, dx, None, i32(d2.shape[0]), i32(d2.shape[1]), f32(p), u64(seed), i32(_lib.F32), stream_ptr())
        return dx.view(dy.shape), None, None


def flat_dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return FlatDropoutFn.apply(x, p, ops.next_seed())


class CrossLayerFn(torch.autograd.Function):
    """layer(s1, s2, mask2) of BertCrossAttentionLayer: q from s1 (B, L1, D), keys / values from s2 (B, L2, D), post-LN.
    Returns the layer output (B, L1, D): in the activation dtype, or -- cfg[6] -- as a fresh fp32 tensor (what the reference's
    LayerNorm yields under autocast; its caller writes into it in place).  s2 may be s1 (self-attention: the RoBERTa layers
    of models/chemberta.py)."""

    @staticmethod
    def forward(ctx, s1, s2, mask2, wq, bq, wk, bk, wv, bv, wo, bo, ln1_w, ln1_b, w1, b1, w2, b2, ln2_w, ln2_b, cfg):
        H, p_attn, p_hid, seeds, dt, eps = cfg[:6]
        out_f32 = len(cfg) > 6 and cfg[6]            # fp32 copy of the output, made here (a fresh tensor: callers may write into it)
        _lib.require_cuda(s1, s2)
        B, L1, D = s1.shape
        L2 = s2.shape[1]
        R1, R2 = B * L1, B * L2
        F_ = w1.shape[0]
        code = DTYPE_CODE[dt]
        sp = stream_ptr()
        scale = 1.0 / math.sqrt(D // H)
        fused = dt == torch.bfloat16 and ops_gemm.supported(D, F_) and (D // H) in (32, 64)
        if dt == torch.bfloat16 and not fused:
            raise _lib.MMDTIError("cross layer: bf16 mode needs hidden %% 64 == 0, hidden <= 512 and head_dim 32 or 64 (got D=%d, H=%d)" % (D, H))
        def both(t, rows):
            """(fp32, activation-dtype) copies of an input stream, each made at most once by the vectorised convert kernel"""
            t = t.detach().reshape(rows, D).contiguous()
            if t.dtype == torch.float32:
                return t, _convert(t, dt)
            if t.dtype == dt:
                return _convert(t, torch.float32), t
            tf = t.float()
            return tf, _convert(tf, dt)

        s1f, s1l = both(s1, R1)
        s2f, s2l = (s1f, s1l) if s2 is s1 else both(s2, R2)
        mask2 = mask2.detach().to(torch.uint8).contiguous()
        lw = lambda t: t.detach().to(dt).contiguous()
        wq_l, wo_l, w1_l, w2_l = lw(wq), lw(wo), lw(w1), lw(w2)
        wkv_l = torch.cat([wk.detach(), wv.detach()], 0).to(dt)
        bkv_l = torch.cat([bk.detach(), bv.detach()], 0).to(dt)
        ln1_wd, ln1_bd, ln2_wd, ln2_bd = (t.detach().float().contiguous() for t in (ln1_w, ln1_b, ln2_w, ln2_b))
        if fused:
            q = ops_gemm.gemm_bias(s1l, wq_l, lw(bq))
            kv = ops_gemm.gemm_bias(s2l, wkv_l, bkv_l)
        else:
            q = torch.addmm(lw(bq), s1l, wq_l.t())
            kv = torch.addmm(bkv_l, s2l, wkv_l.t())
        o, lse = _attn_fwd(q, kv, D, mask2, B, H, L1, L2, scale, p_attn, seeds[0])
        if fused:
            xo1, a, st1 = ops_gemm.gemm_dropres_ln(o, wo_l, lw(bo), s1f, ln1_wd, ln1_bd, p_hid, seeds[1], eps=eps)
            a_res = _convert(a, torch.float32)
            z, u = ops_gemm.gemm_bias_gelu(a, w1_l, lw(b1), store_grad=True)
            xo2, y, st2 = ops_gemm.gemm_dropres_ln(u, w2_l, lw(b2), a_res, ln2_wd, ln2_bd, p_hid, seeds[2], eps=eps)
        else:
            def dropres_ln(t, res, w, b, seed):
                xo = torch.empty((R1, D), device=t.device, dtype=torch.float32)
                yy = torch.empty((R1, D), device=t.device, dtype=dt)
                st = torch.empty((2, R1), device=t.device, dtype=torch.float32)
                call("mmdti_dropres_layernorm_fwd", res, t, xo, w, b, yy, st[0], st[1], i32(R1), i32(D), f32(eps), f32(p_hid), u64(seed),
                     i32(code), i32(code), sp)
                return xo, yy, st
            xo1, a, st1 = dropres_ln(torch.addmm(lw(bo), o, wo_l.t()), s1f, ln1_wd, ln1_bd, seeds[1])
            a_res = a.float()
            z = torch.addmm(lw(b1), a, w1_l.t())
            u = torch.empty_like(z)
            call("mmdti_gelu_fwd", z, u, i64(z.numel()), i32(code), sp)
            xo2, y, st2 = dropres_ln(torch.addmm(lw(b2), u, w2_l.t()), a_res, ln2_wd, ln2_bd, seeds[2])
        ctx.save_for_backward(s1l, s2l, mask2, q, kv, o, lse, xo1, st1, a, z, u, xo2, st2, ln1_wd, ln2_wd, wq_l, wkv_l, wo_l, w1_l, w2_l)
        ctx.cfg = cfg
        ctx.fused = fused
        ctx.dims = (B, L1, L2, D, F_)
        if out_f32:
            return _convert(y, torch.float32, shape=(B, L1, D), copy=True)
        return y.view(B, L1, D)

    @staticmethod
    def backward(ctx, dy):
        (s1l, s2l, mask2, q, kv, o, lse, xo1, st1, a, z, u, xo2, st2, ln1_w, ln2_w, wq_l, wkv_l, wo_l, w1_l, w2_l) = ctx.saved_tensors
        H, p_attn, p_hid, seeds, dt, eps = ctx.cfg[:6]
        B, L1, L2, D, F_ = ctx.dims
        R1, R2 = B * L1, B * L2
        code = DTYPE_CODE[dt]
        sp = stream_ptr()
        dev = dy.device
        fused = ctx.fused
        scale = 1.0 / math.sqrt(D // H)
        dy = dy.reshape(R1, D).contiguous().float()
        red = torch.zeros(9 * D + F_, device=dev, dtype=torch.float32)
        dw_ln1, db_ln1, dw_ln2, db_ln2 = red[0:D], red[D:2 * D], red[2 * D:3 * D], red[3 * D:4 * D]
        db_o, db_2, db_q, db_kv = red[4 * D:5 * D], red[5 * D:6 * D], red[6 * D:7 * D], red[7 * D:9 * D]
        db_1 = red[9 * D:]
        if fused:
            wgrad_ = ops_gemm.gemm_wgrad
            dgrad = ops_gemm.gemm_dgrad
        else:
            wgrad_ = lambda g, x: torch.mm(g.t(), x).float()
            dgrad = lambda g, w: torch.mm(g, w)
        # weight gradients and bias column sums are off the critical path: side stream (parallel branches of a captured graph),
        # joined before the function returns -- same scheme as ops.EncoderLayerFn
        main = torch.cuda.current_stream(dev)
        side = ops._side_stream(dev) if ops.overlap_wgrad else None

        def off_path(fn):
            if side is None:
                return fn()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                out = fn()
            out.record_stream(main)
            return out

        wgrad = lambda g, x: off_path(lambda: wgrad_(g, x))
        # ---- output block: LayerNorm-2 backward + dropout backward of the fc2 output
        dxo2 = torch.empty((R1, D), device=dev, dtype=torch.float32)
        df = torch.empty((R1, D), device=dev, dtype=dt)
        call("mmdti_layernorm_bwd_dropout", _convert(dy, dt), xo2, ln2_w, st2[0], st2[1], None, dxo2, dw_ln2, db_ln2,
             df, db_2, i32(R1), i32(D), f32(p_hid), u64(seeds[2]), i32(code), sp)
        dW2 = wgrad(df, u)
        if fused:
            dz = ops_gemm.gemm_dgrad_gelu(df, w2_l, z, db_1, z_is_grad=True)
        else:
            du = torch.mm(df, w2_l)
            dz = torch.empty_like(z)
            call("mmdti_gelu_bwd", du, z, dz, db_1, i32(R1), i32(F_), i32(code), sp)
        dW1 = wgrad(dz, a)
        # ---- attention block: the gradient at a = LayerNorm-1 output is (dz W1) + dxo2 (residual of the output block);
        # LayerNorm' is linear in its incoming gradient, so the residual part goes first and the GEMM epilogue adds the rest
        t = torch.empty((R1, D), device=dev, dtype=torch.float32)
        call("mmdti_layernorm_bwd", dxo2, xo1, ln1_w, st1[0], st1[1], None, t, dw_ln1, db_ln1, i32(R1), i32(D), i32(_lib.F32), sp)
        if fused:
            dxo1, da = ops_gemm.gemm_dgrad_lnbwd(dz, w1_l, xo1, st1, ln1_w, t, dw_ln1, db_ln1, db_o, p_hid, seeds[1])
        else:
            dh = torch.mm(dz, w1_l)
            dxo1 = torch.empty((R1, D), device=dev, dtype=torch.float32)
            da = torch.empty((R1, D), device=dev, dtype=dt)
            call("mmdti_layernorm_bwd_dropout", dh, xo1, ln1_w, st1[0], st1[1], t, dxo1, dw_ln1, db_ln1, da, db_o, i32(R1), i32(D), f32(p_hid),
                 u64(seeds[1]), i32(code), sp)
        dWo = wgrad(da, o)
        d_o = dgrad(da, wo_l)
        dq, dkv = _attn_bwd(q, kv, D, mask2, o, d_o, lse, B, H, L1, L2, scale, p_attn, seeds[0])
        def qkv_grads():
            sps = stream_ptr()
            call("mmdti_colsum", dq, db_q, i32(R1), i32(D), i32(code), sps)
            call("mmdti_colsum", dkv, db_kv, i32(R2), i32(2 * D), i32(code), sps)
            return wgrad_(dq, s1l), wgrad_(dkv, s2l)

        if side is None:
            dWq, dWkv = qkv_grads()
        else:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dWq, dWkv = qkv_grads()
            dWq.record_stream(main)
            dWkv.record_stream(main)
        ds1 = _convert(dgrad(dq, wq_l), torch.float32, add=dxo1)
        ds2 = _convert(dgrad(dkv, wkv_l), torch.float32)
        if side is not None:
            main.wait_stream(side)
        return (ds1.view(B, L1, D), ds2.view(B, L2, D), None, dWq, db_q, dWkv[:D], db_kv[:D], dWkv[D:], db_kv[D:], dWo, db_o, dw_ln1, db_ln1,
                dW1, db_1, dW2, db_2, dw_ln2, db_ln2, None)


class MaskedPoolFn(torch.autograd.Function):
    """models/mm_model.py:572-576: rows outside the masks zeroed, the two sequences concatenated and summed over dim 1,
    divided by the number of valid rows of both."""

    @staticmethod
    def forward(ctx, x1, m1, x2, m2):
        _lib.require_cuda(x1, x2)
        B, L1, D = x1.shape
        L2 = x2.shape[1]
        if x1.dtype != x2.dtype or x1.dtype not in (torch.float32, torch.bfloat16):
            x1, x2 = x1.float(), x2.float()
        x1, x2 = x1.detach().contiguous(), x2.detach().contiguous()
        m1, m2 = m1.detach().to(torch.uint8).contiguous(), m2.detach().to(torch.uint8).contiguous()
        out = torch.empty((B, D), device=x1.device, dtype=torch.float32)
        inv = torch.empty(B, device=x1.device, dtype=torch.float32)
        call("mmdti_masked_pool_fwd", x1, m1, i32(L1), x2, m2, i32(L2), out, inv, i32(B), i32(D), i32(DTYPE_CODE[x1.dtype]), stream_ptr())
        ctx.save_for_backward(m1, m2, inv)
        ctx.dims = (B, L1, L2, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        m1, m2, inv = ctx.saved_tensors
        B, L1, L2, D = ctx.dims
        dx1 = torch.empty((B, L1, D), device=dout.device, dtype=torch.float32)
        dx2 = torch.empty((B, L2, D), device=dout.device, dtype=torch.float32)
        call("mmdti_masked_pool_bwd", dout.contiguous().float(), inv, m1, i32(L1), m2, i32(L2), dx1, dx2, i32(B), i32(D), stream_ptr())
        return dx1, None, dx2, None


def masked_mean_pool(x1, mask1, x2, mask2):
    """(x1 (B, L1, D), mask1 (B, L1) bool, x2 (B, L2, D), mask2 (B, L2) bool) -> (B, D) f32"""
    return MaskedPoolFn.apply(x1, mask1, x2, mask2)
