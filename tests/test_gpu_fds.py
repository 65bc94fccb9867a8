"""GPU parity of K5 (FDS) against the golden fixture generated from the reference's models/fds.py
and utils/util.py (oracle/make_golden.py:gold_fds).  Bin indices are bit-exact; statistics and the
calibrated features agree to 1e-5 (fp32 everywhere)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from emu_dp import run_emulated
from mmdti_b200 import ops_fds
from mmdti_b200.models.fds import FDS
from oracle import restate

STAT_KEYS = ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked")


def _make(g, device, D=16, nb=12):
    m = FDS(feature_dim=D, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb, bucket_start=0,
            start_update=0, start_smooth=1, kernel="gaussian", ks=5, sigma=1, momentum=0.9, device=device)
    m.min_value, m.bin_width = float(g["cfg.min_value"]), float(g["cfg.bin_width"])
    return m.to(device)


def test_fds_constructor_matches_reference_binning_setup():
    """CPU: StandardScaler + 3-sigma clean + min / bin width exactly like models/fds.py:44-57."""
    g = load_golden("fds")
    raw = torch.randn(400, generator=torch.Generator().manual_seed(60)).double() * 1.7 + 0.3
    m = FDS(feature_dim=16, raw_data=raw.numpy(), col_data="expt", using_scale=True, bucket_num=12, ks=5, sigma=1)
    assert abs(float(m.min_value) - float(g["cfg.min_value"])) < 1e-12
    assert abs(float(m.bin_width) - float(g["cfg.bin_width"])) < 1e-12
    assert rel_err(m.kernel_window, g["cfg.window"]) < 1e-6
    assert sorted(m.state_dict()) == sorted(["epoch", "running_mean", "running_var", "running_mean_last_epoch",
                                             "running_var_last_epoch", "smoothed_mean_last_epoch",
                                             "smoothed_var_last_epoch", "num_samples_tracked"])


@pytest.mark.gpu
def test_fds_golden_flow(report):
    g = load_golden("fds")
    m = _make(g, "cuda")
    labels = g["in.labels"].cuda()
    bins, present = m._bins(labels)
    assert torch.equal(bins.cpu().long(), g["out.bins"].long())                      # bit-exact
    m.update_running_stats(g["in.feats_e0"].cuda(), labels, 0)
    m.update_last_epoch_stats(1)
    for k in STAT_KEYS:
        e = rel_err(getattr(m, k), g["out.e1." + k])
        report("fds e1", k, "%.2e" % e)
        assert e < 1e-5, k
    assert m.running_mean_last_epoch is m.running_mean                                # aliasing kept (Q9)
    x = g["in.smooth_x"].cuda().requires_grad_(True)
    xin = x * 1.0
    xs = m.smooth(xin, labels, 1)
    assert xs.data_ptr() == xin.data_ptr()                                            # in place (Q11)
    (xs * g["in.smooth_up"].cuda()).sum().backward()
    assert rel_err(xs, g["out.smooth"]) < 1e-5
    assert rel_err(x.grad, g["grad.smooth_x"]) < 1e-5
    sub = g["in.sub"].bool()
    xs2 = m.smooth(g["in.smooth_x"][sub].cuda().clone(), labels[sub.cuda()], 1)
    assert rel_err(xs2, g["out.smooth_noedge"]) < 1e-5                                 # tails untouched without edge bins
    assert torch.equal(m.smooth(x.detach().clone(), labels, 0).cpu(), x.detach().cpu())   # epoch < start_smooth
    m.update_running_stats(g["in.feats_e1"].cuda(), labels, 1)
    m.update_last_epoch_stats(2)
    for k in STAT_KEYS:
        e = rel_err(getattr(m, k), g["out.e2." + k])
        report("fds e2", k, "%.2e" % e)
        assert e < 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("N,D,nb", [(5000, 512, 30), (1237, 96, 100), (64, 512, 30)])
def test_fds_vs_oracle_large(N, D, nb):
    gen = torch.Generator().manual_seed(N)
    labels = torch.randn(N, 1, generator=gen) * 1.3
    feats = [torch.randn(N, D, generator=gen) * (1 + i) + 0.1 * i for i in range(2)]
    cfg = dict(min_value=-2.5, bin_width=5.0 / nb, bucket_num=nb, bucket_start=0, start_update=0, start_smooth=1, momentum=0.9)
    win = restate.fds_kernel_window("gaussian", 5, 2)
    st = {"epoch": torch.zeros(1), "running_mean": torch.zeros(nb, D), "running_var": torch.ones(nb, D),
          "running_mean_last_epoch": torch.zeros(nb, D), "running_var_last_epoch": torch.ones(nb, D),
          "smoothed_mean_last_epoch": torch.zeros(nb, D), "smoothed_var_last_epoch": torch.ones(nb, D),
          "num_samples_tracked": torch.zeros(nb)}
    m = FDS(feature_dim=D, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb, ks=5, sigma=2).cuda()
    m.min_value, m.bin_width = cfg["min_value"], cfg["bin_width"]
    for ep in range(2):
        restate.fds_update_running_stats(feats[ep], labels, ep, st, cfg)
        restate.fds_update_last_epoch_stats(ep + 1, st, win)
        m.update_running_stats(feats[ep].cuda(), labels.cuda(), ep)
        m.update_last_epoch_stats(ep + 1)
        for k in STAT_KEYS:
            assert rel_err(getattr(m, k), st[k]) < 2e-5, (ep, k)
        x = torch.randn(N, D, generator=gen)
        want = restate.fds_smooth(x.clone(), labels, ep + 1, st, cfg)
        got = m.smooth(x.cuda().clone(), labels.cuda(), ep + 1)
        assert rel_err(got, want) < 2e-5


@pytest.mark.gpu
def test_fds_data_parallel_emulation():
    """W emulated ranks with a shard of the epoch's features each end up with the single-process buffers."""
    N, D, nb, W = 4096, 128, 20, 4
    gen = torch.Generator().manual_seed(11)
    labels = torch.randn(N, generator=gen)
    feats = torch.randn(N, D, generator=gen) * 2 + 0.5

    def mk():
        m = FDS(feature_dim=D, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb).cuda()
        m.min_value, m.bin_width = -2.0, 4.0 / nb
        return m

    ref = mk()
    ref.update_running_stats(feats.cuda(), labels.cuda(), 0)
    M = N // W

    def ranked(dp, r):
        m = mk()
        m.dp = dp
        sl = slice(r * M, r * M + M)
        m.update_running_stats(feats[sl].cuda(), labels[sl].cuda(), 0)
        return m

    for m in run_emulated(W, ranked, sweeps=4):      # present -> {count,sum} -> m2: three dependent exchanges
        for k in ("running_mean", "running_var", "num_samples_tracked"):
            assert rel_err(getattr(m, k), getattr(ref, k)) < 1e-5, k
