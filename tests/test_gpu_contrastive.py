"""GPU parity of K3 (InfoNCE) and K4 (SupCon / ConR / multi-label) through the C ABI against the
golden fixtures generated from the reference's own models/infonce.py and models/contrastive.py
(oracle/make_golden.py) and against the oracle restatement at larger sizes.

Tolerances (tests/tolerances.py).  fp32 validation mode: loss 5e-6, gradients 1e-5 (relative to the max |reference|).
bf16 tensor-core mode: operands are bf16 unit vectors (2^-9 relative rounding), logits are scaled by
1/t = 10..14.3 before the exponential, so loss 1e-3 (north_star's bf16 bound: met) and gradients 1.2e-2 in the norm sense (observed 4.4e-3: not met)."""
import pytest
import torch

import mmdti_b200
from conftest import load_golden, norm_err, rel_err
from emu_dp import run_emulated
from mmdti_b200 import ops_sim
from mmdti_b200.models import contrastive as ctm
from mmdti_b200.models import infonce as infm
from oracle import restate

pytestmark = pytest.mark.gpu
from tolerances import TOL as _T

MODES = [("fp32",) + _T["sim.fp32"], ("bf16",) + _T["sim.bf16"]]


def _tol(a, b, mode):
    return rel_err(a, b) if mode == "fp32" else norm_err(a, b)


@pytest.mark.parametrize("mode,ltol,gtol", MODES)
@pytest.mark.parametrize("tag", ["n16", "n37", "n64d512"])
def test_infonce_golden(tag, mode, ltol, gtol, report):
    g = load_golden("infonce_" + tag)
    q = g["in.q"].cuda().requires_grad_(True)
    k = g["in.k"].cuda().requires_grad_(True)
    with mmdti_b200.precision(act=mode):
        loss = infm.info_nce(q, k, temperature=0.1)
        loss.backward()
    e = (rel_err(loss, g["out.loss"]), _tol(q.grad, g["grad.q"], mode), _tol(k.grad, g["grad.k"], mode))
    report("infonce", tag, mode, *("%.2e" % x for x in e))
    assert e[0] < ltol and e[1] < gtol and e[2] < gtol, e


@pytest.mark.parametrize("mode,ltol,gtol", MODES)
@pytest.mark.parametrize("n", [16, 45])
def test_ct_golden(n, mode, ltol, gtol, report):
    g = load_golden("ct_n%d" % n)
    f = g["in.feature"].cuda()
    y, yhat, wts = g["in.y"].cuda(), g["in.yhat"].cuda(), g["in.weights"].cuda()
    cases = {
        "regress_w": lambda x: ctm.CT_Regress(x, y, yhat, weights=wts, w=0.2),
        "regress_now": lambda x: ctm.CT_Regress(x, y, yhat, w=0.2),
        "single": lambda x: ctm.CT_Single(x, g["in.cls"].cuda(), yhat),
        "multi": lambda x: ctm.CT_Multi(x, g["in.multi"].cuda(), yhat),
    }
    for name, fn in cases.items():
        x = f.clone().requires_grad_(True)
        with mmdti_b200.precision(act=mode):
            loss = fn(x)
            loss.backward()
        e = (rel_err(loss, g["out." + name]), _tol(x.grad, g["grad." + name], mode))
        report("ct", n, name, mode, *("%.2e" % v for v in e))
        assert e[0] < ltol and e[1] < gtol, (name, e)


@pytest.mark.parametrize("n", [16, 45])
def test_ct_masks_bit_exact(n):
    g = load_golden("ct_n%d" % n)
    for mode, code, lab in (("regress", ops_sim.REGRESS, "in.y"), ("single", ops_sim.SINGLE, "in.cls"),
                            ("multi", ops_sim.MULTI, "in.multi")):
        pos, neg = ops_sim.ct_masks(code, g[lab].cuda(), g["in.yhat"].cuda(), w=0.2)
        assert torch.equal(pos.cpu(), g["out.%s_pos" % mode].bool()), mode
        assert torch.equal(neg.cpu(), g["out.%s_neg" % mode].bool()), mode


def _rand_case(N, D, seed):
    gen = torch.Generator().manual_seed(seed)
    f = torch.randn(N, D, generator=gen)
    y = torch.randn(N, 1, generator=gen)
    yhat = y + 0.3 * torch.randn(N, 1, generator=gen)
    w = torch.rand(N, generator=gen) + 0.5
    cls = torch.randint(0, 10, (N, 1), generator=gen)
    return f, y, yhat, w / w.mean(), cls


@pytest.mark.parametrize("mode,ltol,gtol", MODES)
@pytest.mark.parametrize("N,D", [(200, 512), (517, 50), (1000, 300), (2048, 512)])
def test_losses_vs_oracle_ragged_sizes(N, D, mode, ltol, gtol, report):
    """sizes that are not multiples of the 128-row / 32..128-key tiles, D not a multiple of 64"""
    f, y, yhat, w, cls = _rand_case(N, D, 7 + N)
    f2 = torch.randn(N, D, generator=torch.Generator().manual_seed(N))
    cases = {
        "infonce": (lambda a, b: restate.info_nce(a, b, 0.1), lambda a, b: infm.info_nce(a, b, temperature=0.1)),
        "conr": (lambda a, b: restate.ct_regress(a, y, yhat, weights=w, w=0.2), None),
        "supcon": (lambda a, b: restate.ct_single(a, cls), None),
    }
    for name, (ref_fn, _) in cases.items():
        a = f.clone().requires_grad_(True)
        b = (0.5 * f + f2).clone().requires_grad_(True)
        ref = ref_fn(a, b)
        ref.backward()
        ac, bc = f.cuda().requires_grad_(True), (0.5 * f + f2).cuda().requires_grad_(True)
        with mmdti_b200.precision(act=mode):
            if name == "infonce":
                loss = infm.info_nce(ac, bc, temperature=0.1)
            elif name == "conr":
                loss = ctm.CT_Regress(ac, y.cuda(), yhat.cuda(), weights=w.cuda(), w=0.2)
            else:
                loss = ctm.CT_Single(ac, cls.cuda(), None)
            loss.backward()
        e = [rel_err(loss, ref), _tol(ac.grad, a.grad, mode)]
        if name == "infonce":
            e.append(_tol(bc.grad, b.grad, mode))
        report("sim-vs-oracle", name, N, D, mode, *("%.2e" % v for v in e))
        assert e[0] < ltol and all(v < gtol for v in e[1:]), (name, e)


@pytest.mark.parametrize("N,D", [(4096, 512), (4396, 136)])
def test_large_n_tensor_core_tiles_match_fp32_kernels(N, D, report):
    """N >= 4096 switches phase 1 to 256-key tiles (sim_tc.cu); the fp32 validation kernels, themselves checked against
    the oracle at small N above, are the reference here (the oracle would need an N x N fp32 matrix per loss)"""
    f, y, yhat, w, cls = _rand_case(N, D, 11 + N)
    f2 = torch.randn(N, D, generator=torch.Generator().manual_seed(N))
    out = {}
    for mode in ("fp32", "bf16"):
        res = []
        for name in ("infonce", "conr", "supcon"):
            ac, bc = f.cuda().requires_grad_(True), (0.5 * f + f2).cuda().requires_grad_(True)
            with mmdti_b200.precision(act=mode):
                if name == "infonce":
                    loss = infm.info_nce(ac, bc, temperature=0.1)
                elif name == "conr":
                    loss = ctm.CT_Regress(ac, y.cuda(), yhat.cuda(), weights=w.cuda(), w=0.2)
                else:
                    loss = ctm.CT_Single(ac, cls.cuda(), None)
                loss.backward()
            res.append((name, loss.detach(), ac.grad.clone(), bc.grad.clone() if name == "infonce" else None))
        out[mode] = res
    for (name, l32, g32, gb32), (_, l16, g16, gb16) in zip(out["fp32"], out["bf16"]):
        e = [rel_err(l16, l32), norm_err(g16, g32)] + ([norm_err(gb16, gb32)] if gb32 is not None else [])
        report("sim-large-n", name, N, D, *("%.2e" % v for v in e))
        assert e[0] < _T["sim.bf16"][0] and all(v < _T["sim.bf16"][1] for v in e[1:]), (name, e)      # observed 1.1e-5 / 3.8e-3


def test_fused_and_two_step_backward_agree(monkeypatch):
    """Dp > 256: the fused phase-2 kernel (dA accumulated in TMEM) and the two-step form (tcgen05 coefficient kernel +
    plain GEMM) evaluate the same H_ij in bf16; their gradients agree to bf16 accumulation noise"""
    N, D = 3000, 512
    f, y, yhat, w, cls = _rand_case(N, D, 99)
    grads = {}
    for form in ("fused", "twostep"):
        monkeypatch.setenv("MMDTI_SIM_BWD", form)
        a = f.cuda().requires_grad_(True)
        with mmdti_b200.precision(act="bf16"):
            (ctm.CT_Regress(a, y.cuda(), yhat.cuda(), weights=w.cuda(), w=0.2) + infm.info_nce(a, a.detach().roll(1, 0))).backward()
        grads[form] = a.grad.clone()
    assert norm_err(grads["twostep"], grads["fused"]) < 2e-3


def test_infonce_value_errors():
    for bad in ((torch.randn(4), torch.randn(4, 3)), (torch.randn(4, 3), torch.randn(5, 3)),
                (torch.randn(4, 3), torch.randn(4, 2))):
        with pytest.raises(ValueError):
            infm.info_nce(bad[0].cuda(), bad[1].cuda())


def test_infonce_module_matches_oracle_head():
    torch.manual_seed(3)
    mod = infm.InfoNCE(512, 512).cuda()
    mod.eval()
    gen = torch.Generator().manual_seed(4)
    query, pos = torch.randn(6, 7, 512, generator=gen), torch.randn(6, 9, 512, generator=gen)
    p = {"infonce." + k: v.detach().cpu() for k, v in mod.state_dict().items()}
    ref = restate.infonce_head(query, pos, p)
    with mmdti_b200.precision(act="fp32"):
        got = mod(query.cuda(), pos.cuda())
    assert rel_err(got, ref) < 2e-5
    with mmdti_b200.precision(act="bf16"):            # projection GEMMs on bf16 operands, fp32 accumulate
        got = mod(query.cuda(), pos.cuda())
    assert rel_err(got, ref) < 2e-3, rel_err(got, ref)
    assert sorted(mod.state_dict()) == sorted(["info_proj_query.0.weight", "info_proj_query.0.bias", "info_proj_query.2.weight",
                                               "info_proj_query.2.bias", "info_proj_positive.0.weight",
                                               "info_proj_positive.0.bias", "info_proj_positive.2.weight",
                                               "info_proj_positive.2.bias"])


@pytest.mark.parametrize("mode,ltol,gtol", MODES)
def test_data_parallel_emulation_equals_single_process(mode, ltol, gtol):
    """W = 4 emulated ranks, each owning N/W rows: the loss equals the single-process loss on the
    concatenated batch and every rank's gradient equals its rows of the single-process gradient."""
    N, D, W = 512, 512, 4
    f, y, yhat, w, cls = _rand_case(N, D, 99)
    f2 = 0.5 * f + torch.randn(N, 50 if False else D, generator=torch.Generator().manual_seed(5))
    M = N // W
    with mmdti_b200.precision(act=mode):
        def single():
            a, b = f.cuda().requires_grad_(True), f2.cuda().requires_grad_(True)
            l1 = infm.info_nce(a, b)
            l2 = ctm.CT_Regress(a, y.cuda(), yhat.cuda(), weights=w.cuda())
            l3 = ctm.CT_Single(a, cls.cuda(), None)
            (l1 + l2 + l3).backward()
            return torch.stack([l1, l2, l3]).detach(), a.grad, b.grad

        def ranked(dp, r):
            sl = slice(r * M, r * M + M)
            a, b = f[sl].cuda().requires_grad_(True), f2[sl].cuda().requires_grad_(True)
            l1 = infm.info_nce(a, b, dp=dp)
            l2 = ctm.CT_Regress(a, y[sl].cuda(), yhat[sl].cuda(), weights=w[sl].cuda(), dp=dp)
            l3 = ctm.CT_Single(a, cls[sl].cuda(), None, dp=dp)
            (l1 + l2 + l3).backward()
            return torch.stack([l1, l2, l3]).detach(), a.grad, b.grad

        ls, ga, gb = single()
        outs = run_emulated(W, ranked)
    for r, (lr, gar, gbr) in enumerate(outs):
        sl = slice(r * M, r * M + M)
        assert rel_err(lr, ls) < 1e-5, (r, lr, ls)                      # same kernels, same tiles -> near identical
        assert norm_err(gar, ga[sl]) < 1e-4 and norm_err(gbr, gb[sl]) < 1e-4
