"""K3 / K4 host side: InfoNCE and the SupCon / ConR / multi-label contrastive losses on the
two-phase similarity engine of the C ABI (include/mmdti_b200.h, "contrastive similarity").

    forward : F.normalize -> phase 1 (row statistics over the similarity tiles) -> finalize
    backward: phase 2 (recompute tiles, gradient coefficients, dA) -> normalisation backward

The N x N similarity matrix never exists in the forward; the backward of large problems (see _two_step) materialises
bounded row blocks of the bf16 gradient coefficients H and finishes with one plain GEMM.  Two arithmetic modes (config.fp32_mode()):
  bf16 (production)  bf16 unit vectors on the tcgen05 tensor cores, fp32 accumulation/statistics
  fp32 (validation)  plain fp32 FMA kernels (1e-5 parity with the reference)

Data parallel: pass ``dp=DataParallelCtx``; every rank owns M = N / world anchor rows of the
all-gathered batch (rows [rank*M, rank*M + M)), evaluates them against the GLOBAL keys and gets
the exact gradient of the global loss w.r.t. its own rows (see dist.py).  The returned loss is the
global loss on every rank.

Reference: models/infonce.py:42-98, models/contrastive.py:3-169; closed forms SURVEY.md Appendix B."""
import os

import torch

from . import _lib, config
from ._lib import call, f32, i32, i64, stream_ptr

INFONCE, REGRESS, SINGLE, MULTI = 0, 1, 2, 3
NSTAT = 8
EPS = 1e-12                     # F.normalize default


def _pad64(d):
    return (d + 63) // 64 * 64


def tc_supported(D):
    return _pad64(D) <= 512


class _Operand:
    """row-normalised matrix in the formats the engine consumes"""
    __slots__ = ("f32", "bf16", "inv_norm", "N", "D", "Dp")


def normalize_rows(x, want_f32, want_bf16):
    """x (N,D) f32 -> _Operand (F.normalize(dim=-1))."""
    N, D = x.shape
    op = _Operand()
    op.N, op.D, op.Dp = N, D, _pad64(D)
    op.f32 = torch.empty((N, D), device=x.device, dtype=torch.float32) if want_f32 else None
    op.bf16 = torch.empty((N, op.Dp), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    op.inv_norm = torch.empty(N, device=x.device, dtype=torch.float32)
    call("mmdti_rownorm_fwd", x, i64(x.stride(0)), op.f32, op.bf16, i32(op.Dp), op.inv_norm, i32(N), i32(D), f32(EPS),
         stream_ptr())
    return op


def _label_args(mode, lab):
    """-> (y, yhat, w_thr, e_push, key, C, coef, wrow, wcol) in C-ABI order"""
    return (lab.get("y"), lab.get("yhat"), f32(lab.get("w", 0.0)), f32(lab.get("e", 0.0)), lab.get("key"),
            i32(lab.get("C", 0)), f32(lab.get("coef", 1.0)), lab.get("wrow"), lab.get("wcol"))


def sim_stats(mode, A, B, row_offset, temperature, lab, use_tc):
    """phase 1 for the local anchor rows A against the keys B -> stats (M, 8) f32"""
    M, N = A.N, B.N
    stats = torch.empty((M, NSTAT), device=B.inv_norm.device, dtype=torch.float32)
    if use_tc:
        call("mmdti_sim_stats_tc", A.bf16, B.bf16, i32(M), i32(N), i32(A.Dp), i32(row_offset), i32(mode), f32(temperature),
             *_label_args(mode, lab), stats, stream_ptr())
    else:
        call("mmdti_sim_stats_f32", A.f32, B.f32, i32(M), i32(N), i32(A.D), i32(row_offset), i32(mode), f32(temperature),
             *_label_args(mode, lab), stats, stream_ptr())
    return stats


H_BLOCK_BYTES = 512 << 20        # bound of the bf16 coefficient block of the two-step backward


def _two_step(A, B):
    """Two-step phase 2 (coefficient tiles to HBM, then one plain GEMM) or the fused kernel?  The fused kernel keeps dA
    in TMEM: at Dp > 256 its 512 columns force two 256-column slices (each re-evaluates all H_ij) and the 128 KB anchor
    stripe leaves shared memory for 32-key tiles only, so the two-step form is ~3x faster once the N x N coefficient work
    dominates the launch overheads (measured: profiles/r1_sim_bwd_fused_vs_twostep.log); at Dp <= 256 the fused kernel
    wins.  MMDTI_SIM_BWD=fused|twostep overrides."""
    forced = os.environ.get("MMDTI_SIM_BWD", "auto")
    if forced in ("fused", "twostep"):
        return forced == "twostep"
    return A.Dp > 256 and A.N * B.N >= (1 << 23)


def _sim_grad_two_step(mode, A, B, row_offset, temperature, lab, rs_row, rs_col):
    M, N = A.N, B.N
    dev = B.inv_norm.device
    ldh = (N + 7) // 8 * 8
    rb = max(128, min((M + 127) // 128 * 128, H_BLOCK_BYTES // (2 * ldh) // 128 * 128))
    hbuf = torch.empty((min(rb, M), ldh), device=dev, dtype=torch.bfloat16)
    out = None
    for r0 in range(0, M, rb):
        m = min(rb, M - r0)
        call("mmdti_sim_coef_tc", A.bf16[r0:], B.bf16, i32(m), i32(N), i32(A.Dp), i32(row_offset + r0), i32(mode),
             f32(temperature), *_label_args(mode, lab), rs_row[r0:], rs_col, hbuf, i64(ldh), stream_ptr())
        # dA block = H (m, N) @ B (N, Dp) on the own tcgen05 GEMM (fp32 accumulate and result); columns of H beyond N
        # are never written by the coefficient kernel and are not read here (K = N, row stride ldh)
        blk = torch.empty((m, A.Dp), device=dev, dtype=torch.float32)
        if N % 8 == 0:
            call("mmdti_gemm_nn_f32", hbuf, i64(ldh), B.bf16, i64(A.Dp), blk, i64(A.Dp), i32(m), i32(A.Dp), i32(N), stream_ptr())
        else:
            # N not a multiple of 8: the padding columns [N, ldh) of H are zeroed and the key matrix is padded with zero rows
            hbuf[:m, N:].zero_()
            bpad = torch.zeros((ldh, A.Dp), device=dev, dtype=torch.bfloat16)
            bpad[:N] = B.bf16
            call("mmdti_gemm_nn_f32", hbuf, i64(ldh), bpad, i64(A.Dp), blk, i64(A.Dp), i32(m), i32(A.Dp), i32(ldh), stream_ptr())
        if m == M:
            out = blk
        else:
            if out is None:
                out = torch.empty((M, A.Dp), device=dev, dtype=torch.float32)
            out[r0:r0 + m] = blk
    return out if A.D == A.Dp else out[:, :A.D].contiguous()


def sim_grad(mode, A, B, row_offset, temperature, lab, rs_row, rs_col, use_tc):
    """phase 2: dA (M, D) f32 = sum_j H_ij b_j"""
    M, N = A.N, B.N
    if use_tc and _two_step(A, B):
        return _sim_grad_two_step(mode, A, B, row_offset, temperature, lab, rs_row, rs_col)
    dA = torch.empty((M, A.D), device=B.inv_norm.device, dtype=torch.float32)
    if use_tc:
        call("mmdti_sim_grad_tc", A.bf16, B.bf16, i32(M), i32(N), i32(A.Dp), i32(row_offset), i32(mode), f32(temperature),
             *_label_args(mode, lab), rs_row, rs_col, dA, i64(A.D), stream_ptr())
    else:
        call("mmdti_sim_grad_f32", A.f32, B.f32, i32(M), i32(N), i32(A.D), i32(row_offset), i32(mode), f32(temperature),
             *_label_args(mode, lab), rs_row, rs_col, dA, stream_ptr())
    return dA


def _rownorm_bwd(g, op, dx, scale, gscale, accumulate=False):
    call("mmdti_rownorm_bwd", g, op.f32, op.inv_norm, dx, i64(dx.stride(0)), i32(op.N), i32(op.D), f32(scale), gscale,
         f32(EPS), i32(1 if accumulate else 0), stream_ptr())


def _gather_packed(dp, tensors):
    """ONE all-gather for several per-row tensors (M, ...) of any dtypes: their rows are packed side by side into a byte
    matrix, gathered rank-major, and unpacked again.  Exchange 1 is latency-bound (a few MB over NVSwitch), so the number
    of collectives, not their size, is what the step pays for.  None entries pass through."""
    live = [(i, t) for i, t in enumerate(tensors) if t is not None]
    M = live[0][1].shape[0]
    parts, meta = [], []
    for i, t in live:
        t2 = t.contiguous().view(M, -1)
        b = t2.view(torch.uint8)
        pad = (-b.shape[1]) % 16                      # 16-byte aligned columns: every unpacked view is aligned
        parts.append(b if pad == 0 else torch.nn.functional.pad(b, (0, pad)))
        meta.append((i, t.dtype, tuple(t.shape[1:]), b.shape[1]))
    packed = torch.cat(parts, dim=1)
    gathered = dp.all_gather_rows(packed)
    out = [None] * len(tensors)
    off = 0
    for (i, dt, tail, nbytes), part in zip(meta, parts):
        col = gathered[:, off:off + nbytes].contiguous().view(dt)
        out[i] = col.view((gathered.shape[0],) + tail)
        off += part.shape[1]
    return out


def _gather_operands(ops, dp, extra=()):
    """all-gather normalised operands (and extra per-row tensors) over the data-parallel group in ONE collective."""
    if dp is None or dp.world == 1:
        return list(ops), list(extra)
    flat = []
    for op in ops:
        flat += [op.f32, op.bf16, op.inv_norm]
    got = _gather_packed(dp, flat + list(extra))
    res = []
    for k, op in enumerate(ops):
        g = _Operand()
        g.N, g.D, g.Dp = op.N * dp.world, op.D, op.Dp
        g.f32, g.bf16, g.inv_norm = got[3 * k], got[3 * k + 1], got[3 * k + 2]
        res.append(g)
    return res, got[3 * len(ops):]


def _use_tc(D):
    if config.fp32_mode():
        return False
    if not tc_supported(D):
        raise _lib.MMDTIError("contrastive kernels: feature dim %d > 512 is not supported by the tensor-core path "
                              "(use mmdti_b200.precision(act='fp32') for validation-mode kernels up to 512)" % D)
    return True


# ================================================================================ K3 InfoNCE
class InfoNCEFn(torch.autograd.Function):
    """loss = (CE(q̂ p̂ᵀ/t, arange) + CE(p̂ q̂ᵀ/t, arange)) / 2   (models/infonce.py:70-98)."""

    @staticmethod
    def forward(ctx, query, positive, temperature, reduction, dp):
        _lib.require_cuda(query, positive)
        q = query.detach().float().contiguous()
        p = positive.detach().float().contiguous()
        M, D = q.shape
        use_tc = _use_tc(D)
        world = 1 if dp is None else dp.world
        N = M * world
        off = 0 if dp is None else dp.rank * M
        qo = normalize_rows(q, True, use_tc)
        po = normalize_rows(p, True, use_tc)
        (qg, pg), _ = _gather_operands([qo, po], dp)          # exchange 1: one collective for both modalities
        st_r = sim_stats(INFONCE, qo, pg, off, temperature, {}, use_tc)      # rows of  q̂ p̂ᵀ
        st_c = sim_stats(INFONCE, po, qg, off, temperature, {}, use_tc)      # rows of  p̂ q̂ᵀ  = columns of the above
        coef = (0.5 / N) if reduction == "mean" else 0.5
        lse = torch.empty((2, M), device=q.device, dtype=torch.float32)
        loss = torch.zeros((), device=q.device, dtype=torch.float32)
        call("mmdti_infonce_finalize", st_r, lse[0], loss, i32(M), f32(temperature), f32(coef), stream_ptr())
        call("mmdti_infonce_finalize", st_c, lse[1], loss, i32(M), f32(temperature), f32(coef), stream_ptr())
        if world > 1:
            # one collective for both log-sum-exp vectors and the loss partials (third column, constant per rank)
            got = dp.all_gather_rows(torch.stack([lse[0], lse[1], loss.expand(M)], dim=1).contiguous())
            lse_g = got[:, :2].t().contiguous()
            loss = got[:, 2].view(world, M)[:, 0].sum()
        else:
            lse_g = lse
        ctx.ops = (qo, po, qg, pg)
        ctx.save_for_backward(lse, lse_g)
        ctx.cfg = (temperature, coef, off, use_tc, dp, query.dtype, positive.dtype)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        lse, lse_g = ctx.saved_tensors
        qo, po, qg, pg = ctx.ops
        temperature, coef, off, use_tc, dp, qdt, pdt = ctx.cfg
        gscale = g_loss.detach().float().contiguous()
        scale = coef / temperature * (1.0 if dp is None else dp.grad_scale)
        # dq̂_i = sum_j (softmax_row_i + softmax_col_j - 2 d_ij) p̂_j ;  dp̂ symmetric with the roles swapped
        dqh = sim_grad(INFONCE, qo, pg, off, temperature, {}, lse[0], lse_g[1], use_tc)
        dph = sim_grad(INFONCE, po, qg, off, temperature, {}, lse[1], lse_g[0], use_tc)
        dq = torch.empty_like(dqh)
        dpp = torch.empty_like(dph)
        _rownorm_bwd(dqh, qo, dq, scale, gscale)
        _rownorm_bwd(dph, po, dpp, scale, gscale)
        return dq.to(qdt), dpp.to(pdt), None, None, None


def info_nce_loss(query, positive, temperature=0.1, reduction="mean", dp=None):
    if reduction not in ("mean", "sum"):
        raise _lib.MMDTIError("mmdti_b200 InfoNCE kernel: reduction must be 'mean' or 'sum'")
    return InfoNCEFn.apply(query, positive, float(temperature), reduction, dp)


# ================================================================================ K4 SupCon / ConR / multi
class ContrastiveFn(torch.autograd.Function):
    """CT_Regress / CT_Single / CT_Multi (models/contrastive.py).  Gradient flows into `feature`
    only (labels, predictions and weights only build masks / constants)."""

    @staticmethod
    def forward(ctx, feature, mode, lab_local, temperature, dp):
        _lib.require_cuda(feature)
        f = feature.detach().reshape(feature.shape[0], -1).float().contiguous()
        M, D = f.shape
        use_tc = _use_tc(D)
        world = 1 if dp is None else dp.world
        N = M * world
        off = 0 if dp is None else dp.rank * M
        fo = normalize_rows(f, True, use_tc)
        lab = dict(lab_local)
        names = ("y", "yhat", "key", "wrow", "wcol")
        (fg,), extra = _gather_operands([fo], dp, [lab.get(k) for k in names])      # features + labels: one collective
        if world > 1:
            for k, v in zip(names, extra):
                if v is not None:
                    lab[k] = v
        st = sim_stats(mode, fo, fg, off, temperature, lab, use_tc)
        rowstat = torch.empty((M, 2), device=f.device, dtype=torch.float32)
        loss = torch.zeros((), device=f.device, dtype=torch.float32)
        call("mmdti_ct_finalize", st, rowstat, loss, i32(M), i32(N), i32(mode), f32(lab.get("w", 0.0)), stream_ptr())
        if world > 1:
            got = dp.all_gather_rows(torch.cat([rowstat, loss.expand(M, 1)], dim=1).contiguous())
            rowstat_g = got[:, :2].contiguous()
            loss = got[:, 2].view(world, M)[:, 0].sum()
        else:
            rowstat_g = rowstat
        ctx.ops = (fo, fg, lab)
        ctx.save_for_backward(rowstat, rowstat_g)
        ctx.cfg = (mode, temperature, off, use_tc, dp, feature.dtype, feature.shape)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        rowstat, rowstat_g = ctx.saved_tensors
        fo, fg, lab = ctx.ops
        mode, temperature, off, use_tc, dp, fdt, fshape = ctx.cfg
        gscale = g_loss.detach().float().contiguous()
        scale = 1.0 / temperature * (1.0 if dp is None else dp.grad_scale)
        dfh = sim_grad(mode, fo, fg, off, temperature, lab, rowstat, rowstat_g, use_tc)
        df = torch.empty_like(dfh)
        _rownorm_bwd(dfh, fo, df, scale, gscale)
        return df.to(fdt).reshape(fshape), None, None, None, None


def _mean_rows(t):
    """mean over everything but the batch axis -> (N,) f32 (contrastive.py:8-15)."""
    t = t.detach().reshape(t.shape[0], -1).float()
    return (t[:, 0] if t.shape[1] == 1 else t.mean(1)).contiguous()


def ct_regress(feature, depth, output, weights=None, w=0.2, t=0.07, e=0.01, dp=None):
    lab = {"y": _mean_rows(depth), "yhat": _mean_rows(output), "w": float(w), "e": float(e)}
    if weights is not None:
        lab["wrow"] = _mean_rows(weights)
    return ContrastiveFn.apply(feature, REGRESS, lab, float(t), dp)


def _weights_factors(weights, n, device):
    """CT_Single/CT_Multi pushing weights as passed (contrastive.py:94,97): scalar or (1,) -> constant,
    (N,) -> along columns, (N,1) -> along rows.  Returns (wrow, wcol, constant)."""
    if weights is None:
        return None, None
    wt = torch.as_tensor(weights).detach()
    if wt.numel() == 1 and not wt.is_cuda:
        # the reference's default ``weights=tensor([1])`` lives on the host: read it there (no host-to-device copy, which
        # a CUDA-graph capture would refuse) and fill the constant on the device
        return torch.full((n,), float(wt.reshape(()).item()), device=device, dtype=torch.float32), None
    wt = wt.to(device=device, dtype=torch.float32)
    if wt.numel() == 1:
        c = wt.reshape(1).expand(n).contiguous()
        return c, None
    if wt.dim() == 1 and wt.shape[0] == n:
        return None, wt.contiguous()
    if wt.dim() == 2 and wt.shape == (n, 1):
        return wt[:, 0].contiguous(), None
    raise _lib.MMDTIError("contrastive weights of shape %s are not supported (scalar, (N,) or (N,1))" % (tuple(wt.shape),))


def ct_single(feature, depth, output=None, weights=None, t=0.07, dp=None):
    n = feature.shape[0]
    key = depth.detach().reshape(n, -1)
    if key.shape[1] != 1:
        raise _lib.MMDTIError("CT_Single expects one label per sample (got %d); use CT_Multi" % key.shape[1])
    if key.dtype.is_floating_point:
        # label equality on floats == equality of the float64 bit patterns once -0.0 is folded into +0.0 (NaN labels, which
        # the reference's == never matches, are not supported); no host synchronisation: the step must stay graph-capturable
        key = (key.double() + 0.0).view(torch.int64)
    lab = {"key": key.long().contiguous(), "C": 1}
    lab["wrow"], lab["wcol"] = _weights_factors(weights, n, feature.device)
    return ContrastiveFn.apply(feature, SINGLE, lab, float(t), dp)


def ct_multi(feature, depth, output=None, weights=None, t=0.07, coef=1, dp=None):
    n = feature.shape[0]
    key = depth.detach().reshape(n, -1)
    C = key.shape[1]
    if key.dtype.is_floating_point:
        key = (key.double() + 0.0).view(torch.int64)
    lab = {"key": key.long().contiguous(), "C": C, "coef": float(coef)}
    lab["wrow"], lab["wcol"] = _weights_factors(weights, n, feature.device)
    return ContrastiveFn.apply(feature, MULTI, lab, float(t), dp)


def ct_masks(mode, depth, output=None, w=0.2, coef=1.0):
    """Debug/test export: boolean (pos, neg) masks the kernels use (bit-exact vs the reference)."""
    n = depth.shape[0]
    dev = depth.device
    pos = torch.empty((n, n), device=dev, dtype=torch.uint8)
    neg = torch.empty_like(pos)
    if mode == REGRESS:
        y, yh = _mean_rows(depth), _mean_rows(output)
        call("mmdti_ct_masks", i32(mode), i32(n), y, yh, f32(w), None, i32(0), f32(coef), pos, neg, stream_ptr())
    else:
        key = depth.detach().reshape(n, -1).long().contiguous()
        call("mmdti_ct_masks", i32(mode), i32(n), None, None, f32(w), key, i32(key.shape[1]), f32(coef), pos, neg,
             stream_ptr())
    return pos.bool(), neg.bool()
