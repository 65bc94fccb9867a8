"""K5 host side: FDS binning, calibration (smooth) and running statistics on the C ABI
(include/mmdti_b200.h, "FDS").  No per-sample host sync, no torch.unique loop."""
import torch

from . import _lib
from ._lib import call, f32, i32, i64, stream_ptr


def label_column(labels):
    """labels[:, 0] if 2-D else labels (models/fds.py:120-123,160-163) as a strided f32 view."""
    l0 = labels[:, 0] if labels.dim() > 1 else labels
    l0 = l0.detach()
    if l0.dtype != torch.float32:
        l0 = l0.float()
    return l0


def fds_bin(labels, min_value, bin_width, bucket_start, bucket_num, dp=None):
    """-> (bins (N) int32, present (nb) int32).  Under data parallelism `present` is OR-ed over the
    ranks: the reference's loop runs over the bins of the whole (global) batch."""
    l0 = label_column(labels)
    _lib.require_cuda(l0)
    N = l0.shape[0]
    bins = torch.empty(N, device=l0.device, dtype=torch.int32)
    present = torch.empty(bucket_num - bucket_start, device=l0.device, dtype=torch.int32)
    call("mmdti_fds_bin", l0, i64(l0.stride(0) if N > 1 else 1), i32(N), f32(min_value), f32(bin_width), i32(bucket_start),
         i32(bucket_num), bins, present, stream_ptr())
    if dp is not None and dp.world > 1:
        dp.all_reduce_max_(present)
    return bins, present


class FDSSmoothFn(torch.autograd.Function):
    """FDS.smooth: in-place calibration of `features`; gradient = dy * sqrt(clamp(v2/v1)) on the
    transformed entries (utils/util.py:159-169)."""

    @staticmethod
    def forward(ctx, features, bins, present, m1, v1, m2, v2, bucket_start, bucket_num):
        N, D = features.shape
        if features.dtype != torch.float32 or features.stride(1) != 1:
            raise _lib.MMDTIError("FDS.smooth: features must be f32 with unit column stride (got %s, strides %s)"
                                  % (features.dtype, features.stride()))
        nb = bucket_num - bucket_start
        work = torch.empty(3 * nb * D + nb, device=features.device, dtype=torch.float32)      # factor / mean tables
        call("mmdti_fds_smooth_fwd", features, i64(features.stride(0)), bins, present, i32(N), i32(D), i32(bucket_start),
             i32(bucket_num), m1, v1, m2, v2, work, stream_ptr())
        ctx.mark_dirty(features)
        ctx.save_for_backward(bins, present, v1.clone(), v2.clone())
        ctx.cfg = (N, D, bucket_start, bucket_num)
        return features

    @staticmethod
    def backward(ctx, dy):
        bins, present, v1, v2 = ctx.saved_tensors
        N, D, bs, bn = ctx.cfg
        dy = dy.contiguous().float()
        dx = torch.empty_like(dy)
        work = torch.empty(3 * (bn - bs) * D + (bn - bs), device=dy.device, dtype=torch.float32)
        call("mmdti_fds_smooth_bwd", dy, dx, bins, present, i32(N), i32(D), i32(bs), i32(bn), v1, v2, work, stream_ptr())
        return dx, None, None, None, None, None, None, None, None


def fds_update_running_stats(features, bins, present, bucket_start, bucket_num, running_mean, running_var, tracked,
                             momentum, first_update, dp=None):
    """FDS.update_running_stats (models/fds.py:127-153) for one epoch's features (N, D)."""
    features = features.detach()
    if features.dtype != torch.float32 or features.stride(1) != 1:
        features = features.float().contiguous()
    N, D = features.shape
    nb = bucket_num - bucket_start
    dev = features.device
    seg = torch.empty(nb + 1, device=dev, dtype=torch.int32)
    order = torch.empty(N, device=dev, dtype=torch.int32)
    nbp = (nb + 3) // 4 * 4                                                  # sum1 starts 16-byte aligned (vector kernels)
    acc = torch.zeros(nbp + nb * D, device=dev, dtype=torch.float32)         # [count (padded) | sum1] one buffer, one all-reduce
    count, sum1 = acc[:nb], acc[nbp:].view(nb, D)
    m2 = torch.empty((nb, D), device=dev, dtype=torch.float32)
    sp = stream_ptr()
    call("mmdti_fds_group", bins, present, i32(N), i32(bucket_start), i32(bucket_num), seg, order, count, sp)
    call("mmdti_fds_bucket_sums", features, i64(features.stride(0)), seg, order, sum1, i32(N), i32(D), i32(nb), sp)
    if dp is not None and dp.world > 1:
        acc.copy_(dp.all_reduce_sum(acc))
    call("mmdti_fds_bucket_m2", features, i64(features.stride(0)), seg, order, sum1, count, m2, i32(N), i32(D), i32(nb), sp)
    if dp is not None and dp.world > 1:
        m2 = dp.all_reduce_sum(m2)
    call("mmdti_fds_ema", count, sum1, m2, running_mean, running_var, tracked, i32(nb), i32(D),
         f32(-1.0 if momentum is None else momentum), i32(1 if first_update else 0), sp)
    return count


def fds_window_smooth(src, window):
    """reflect-pad + conv1d along the bucket axis (models/fds.py:90-99) -> new (nb, D) tensor"""
    nb, D = src.shape
    out = torch.empty_like(src)
    call("mmdti_fds_window", src.contiguous(), window, out, i32(nb), i32(D), i32(window.numel()), stream_ptr())
    return out
