"""GPU: edge cases of the hot path against the oracle — extreme sequence lengths, degenerate label sets
(no positives / no negatives), single-sample batches, tile-boundary sizes, unsupported sizes failing loudly,
FDS with empty / out-of-range / single buckets."""
import numpy as np
import pytest
import torch

import mmdti_b200
from conftest import rel_err
from mmdti_b200 import ops
from mmdti_b200._lib import MMDTIError
from mmdti_b200.models.contrastive import CT_Regress, CT_Single
from mmdti_b200.models.fds import FDS
from mmdti_b200.models.infonce import info_nce
from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [1, 2, 8, 264])
def test_pair_attention_extreme_lengths(L):
    B, H = 2, 64
    g = torch.Generator().manual_seed(L)
    qkv = torch.randn(B * L, 3 * 512, generator=g) * 0.5
    bias = torch.randn(B * H, L, L, generator=g)
    d_o = torch.randn(B * L, 512, generator=g)
    with mmdti_b200.precision(act="fp32"):
        q = qkv.cuda().requires_grad_(True)
        b = bias.cuda().requires_grad_(True)
        pad = ops.PairPadFn.apply(b, B, H, L, torch.float32)
        o, s = ops.pair_attention(q, pad, B, H, L, 8 ** -0.5)
        (o * d_o.cuda()).sum().backward()
    qr, br = qkv.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    D = 512

    def heads(t):
        return t.view(B, L, H, 8).transpose(1, 2).reshape(B * H, L, 8)
    o_ref, s_ref = restate.pair_attention(heads(qr[:, :D]), heads(qr[:, D:2 * D]), heads(qr[:, 2 * D:]), br, 8 ** -0.5)
    o_ref = o_ref.view(B, H, L, 8).transpose(1, 2).reshape(B * L, D)
    (o_ref * d_o).sum().backward()
    assert rel_err(o, o_ref) < 1e-5
    assert rel_err(ops.pair_unpad(s, L), s_ref) < 1e-5
    for got, want in ((q.grad, qr.grad), (b.grad, br.grad)):        # L = 1: the softmax gradient is exactly 0 -> absolute tolerance
        assert (got.cpu() - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())


def test_unsupported_sizes_fail_loudly():
    with pytest.raises(MMDTIError):
        ops.pair_ld(265)
    f = torch.randn(8, 600).cuda()
    with pytest.raises(MMDTIError):
        info_nce(f, f)                                              # bf16 tensor-core path: D <= 512
    with mmdti_b200.precision(act="fp32"), pytest.raises(MMDTIError):
        info_nce(f, f)                                              # fp32 kernels: D <= 512
    with pytest.raises(MMDTIError):
        info_nce(torch.randn(4, 8), torch.randn(4, 8))              # CPU tensors: no fallback


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 3e-3)])
@pytest.mark.parametrize("N,D", [(1, 50), (2, 512), (129, 64), (257, 512), (33, 1)])
def test_contrastive_small_and_boundary_sizes(N, D, mode, tol):
    g = torch.Generator().manual_seed(N * 1000 + D)
    f, f2 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    y = torch.randn(N, 1, generator=g)
    yhat = y + 0.3 * torch.randn(N, 1, generator=g)
    cls = torch.randint(0, 3, (N, 1), generator=g)
    ref = [restate.info_nce(f, f2, 0.1), restate.ct_regress(f, y, yhat, w=0.2), restate.ct_single(f, cls)]
    with mmdti_b200.precision(act=mode):
        got = [info_nce(f.cuda(), f2.cuda()), CT_Regress(f.cuda(), y.cuda(), yhat.cuda(), w=0.2), CT_Single(f.cuda(), cls.cuda(), None)]
    for a, b in zip(got, ref):
        assert abs(a.item() - b.item()) <= tol * max(1.0, abs(b.item())), (a.item(), b.item())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_contrastive_degenerate_label_sets(mode):
    N, D = 64, 128
    g = torch.Generator().manual_seed(1)
    f = torch.randn(N, D, generator=g)
    same = torch.zeros(N, 1, dtype=torch.long)                      # one class: no negatives -> every row zeroed
    distinct = torch.arange(N).view(N, 1)                           # all different: no positives
    y = torch.zeros(N, 1)                                           # ConR: every pair is a positive, no negatives
    ywide = torch.arange(N).float().view(N, 1) * 10                 # ConR: no positives; negatives need close predictions
    with mmdti_b200.precision(act=mode):
        for lab in (same, distinct):
            x = f.cuda().requires_grad_(True)
            loss = CT_Single(x, lab.cuda(), None)
            loss.backward()
            want = restate.ct_single(f, lab)
            assert abs(loss.item() - want.item()) < 1e-6 and want.item() == 0.0
            assert x.grad.abs().max().item() == 0.0
        for yy, yh in ((y, y), (ywide, torch.zeros(N, 1))):
            x = f.cuda().requires_grad_(True)
            loss = CT_Regress(x, yy.cuda(), yh.cuda(), w=0.2)
            loss.backward()
            want = restate.ct_regress(f, yy, yh, w=0.2)
            assert abs(loss.item() - want.item()) < 1e-6, (loss.item(), want.item())
            assert torch.isfinite(x.grad).all()


def _fds(nb, D, bucket_start=0, momentum=0.9):
    m = FDS(feature_dim=D, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb,
            bucket_start=bucket_start, momentum=momentum).cuda()
    m.min_value, m.bin_width = 0.0, 1.0
    return m


def _state(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("case", ["out_of_range", "single_bucket", "one_sample", "bucket_start", "momentum_none"])
def test_fds_edge_cases(case):
    nb, D, N = 6, 32, 40
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(N, D, generator=g) + 1
    bs, mom = 0, 0.9
    if case == "out_of_range":
        labels = torch.full((N,), 99.0)
        labels[: N // 2] = -7.0                                      # no edge bin present: nothing is grouped
    elif case == "single_bucket":
        labels = torch.full((N,), 2.5)
    elif case == "one_sample":
        labels, feats = torch.tensor([3.2]), feats[:1]
    elif case == "bucket_start":
        labels, bs = torch.rand(N, generator=g) * 8 - 1, 2
    else:
        labels, mom = torch.rand(N, generator=g) * 6, None
    m = _fds(nb, D, bs, mom)
    cfg = dict(min_value=0.0, bin_width=1.0, bucket_num=nb, bucket_start=bs, start_update=0, start_smooth=1, momentum=mom)
    st = _state(m)
    win = restate.fds_kernel_window("gaussian", 5, 2)
    for ep in range(2):
        x = feats * (1 + ep)
        restate.fds_update_running_stats(x, labels, ep, st, cfg)
        restate.fds_update_last_epoch_stats(ep + 1, st, win)
        m.update_running_stats(x.cuda(), labels.cuda(), ep)
        m.update_last_epoch_stats(ep + 1)
        for k in ("running_mean", "running_var", "num_samples_tracked", "smoothed_mean_last_epoch", "smoothed_var_last_epoch"):
            assert rel_err(getattr(m, k), st[k]) < 1e-5, (case, ep, k)
        z = torch.randn(feats.shape[0], D, generator=g)
        want = restate.fds_smooth(z.clone(), labels, ep + 1, st, cfg)
        got = m.smooth(z.cuda().clone(), labels.cuda(), ep + 1)
        assert rel_err(got, want) < 1e-5, (case, ep)
