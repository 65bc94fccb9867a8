"""GPU parity of the cross-modal fusion block against (i) the fixture made by the reference's own
``CrossAttentionModel`` (oracle/make_golden.py:gold_cross_modal; models/mm_model.py:379-406, models/mm_module.py:493-677)
and (ii) the CPU restatement (oracle/restate.py:cross_layer) with the library's dropout masks replayed.

Tolerances: fp32 validation mode at the 1e-5 class (north_star), bf16 mode per tests/tolerances.py ("cross.*")."""
import math

import pytest
import torch

from conftest import load_golden, norm_err, rel_err
from oracle import restate
from oracle.detw import det_state_dict, det_tensor
from tolerances import TOL

pytestmark = pytest.mark.gpu


def _ref_attn(q, kv, mask2, B, H, Lq, Lk, keep=None, p=0.0):
    D = q.shape[1]
    hd = D // H
    q4 = q.view(B, Lq, H, hd).permute(0, 2, 1, 3)
    k4 = kv[:, :D].reshape(B, Lk, H, hd).permute(0, 2, 1, 3)
    v4 = kv[:, D:].reshape(B, Lk, H, hd).permute(0, 2, 1, 3)
    s = q4 @ k4.transpose(-1, -2) / math.sqrt(hd) + ((1.0 - mask2.to(q.dtype)) * -10000.0)[:, None, None, :]
    pr = torch.softmax(s, -1)
    if keep is not None:
        pr = pr * keep.to(pr.dtype) / (1.0 - p)
    return (pr @ v4).permute(0, 2, 1, 3).reshape(B * Lq, D)


@pytest.mark.parametrize("B,H,Lq,Lk,hd", [(2, 4, 13, 10, 32), (3, 16, 66, 70, 32), (2, 8, 130, 200, 64), (1, 2, 64, 64, 32), (2, 2, 5, 129, 64),
                                          (4100, 16, 2, 3, 32)])      # B * H beyond the 65535 grid-y limit (config 3 on one GPU)
@pytest.mark.parametrize("p", [0.0, 0.2])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cross_attention_core(B, H, Lq, Lk, hd, p, mode, report):
    from mmdti_b200 import ops_cross
    D = H * hd
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    g = torch.Generator().manual_seed(B * 1000 + Lq)
    q = (torch.randn(B * Lq, D, generator=g) * 1.5).cuda().to(dt).requires_grad_(True)
    kv = (torch.randn(B * Lk, 2 * D, generator=g) * 1.5).cuda().to(dt).requires_grad_(True)
    up = torch.randn(B * Lq, D, generator=g).cuda()
    mask2 = torch.ones(B, Lk, dtype=torch.bool)
    for b in range(min(B, 8)):
        mask2[b, max(1, Lk - 3 * b - (Lk // 3 if b else 0)):] = False
    mask2 = mask2.cuda()
    seed = 1234567 + Lk
    keep = ops_cross.cross_attn_dropout_mask(B, H, Lq, Lk, p, seed) if p > 0 else None
    if keep is not None:
        frac = keep.float().mean().item()
        assert abs(frac - (1 - p)) < 0.02, frac
    o = ops_cross.CrossAttnFn.apply(q, kv, mask2.to(torch.uint8), B, H, Lq, Lk, p, seed)
    (o.float() * up).sum().backward()
    q64, kv64 = q.detach().double().requires_grad_(True), kv.detach().double().requires_grad_(True)
    o64 = _ref_attn(q64, kv64, mask2, B, H, Lq, Lk, keep, p)
    (o64 * up.double()).sum().backward()
    errs = dict(o=rel_err(o, o64), dq=rel_err(q.grad, q64.grad), dkv=rel_err(kv.grad, kv64.grad))
    report("cross_attn_core", (B, H, Lq, Lk, hd), p, mode, {k: "%.1e" % v for k, v in errs.items()})
    tol = TOL["cross.attn." + mode]
    assert max(errs.values()) < tol, errs


def _load_dropin(g):
    from mmdti_b200.models.cross_modal import CrossAttentionModel, crossmodal_config
    H, D, Fd, seed, ROWS = [int(v) for v in g["cfg"]]
    net = CrossAttentionModel(crossmodal_config(hidden_size=D, num_attention_heads=H, intermediate_size=Fd), num_layers=1)
    sd = det_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=seed, std=0.05)
    net.load_state_dict(sd, strict=True)
    return net.cuda(), ROWS


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cross_modal_golden(mode, report):
    import mmdti_b200
    from mmdti_b200.models.cross_modal import fuse_and_pool
    g = load_golden("cross_modal")
    net, ROWS = _load_dropin(g)
    net.eval()
    x1, x2 = g["in.x1"].cuda().requires_grad_(True), g["in.x2"].cuda().requires_grad_(True)
    m1, m2 = g["in.m1"].cuda(), g["in.m2"].cuda()
    with mmdti_b200.precision(act=mode, pair=mode):
        t2g, g2t = net(x1, x2, m1, m2)
        pooled = fuse_and_pool(t2g, g2t, m1, m2)
        (pooled * g["in.up"].cuda()).sum().backward()
    # rows outside the masks are zeroed by the caller (models/mm_model.py:572-573): compare the valid rows
    errs = dict(t2g=rel_err(t2g.float()[m1], g["out.t2g"].cuda()[m1]), g2t=rel_err(g2t.float()[m2], g["out.g2t"].cuda()[m2]),
                pooled=rel_err(pooled, g["out.pooled"]), dx1=norm_err(x1.grad, g["grad.x1"]), dx2=norm_err(x2.grad, g["grad.x2"]))
    named = dict(net.named_parameters())
    gerr = {}
    bias_scale = g["grad.graph_attention.layer.0.attention.self.query.bias"].abs().max().item()
    for k, v in g.items():
        if not k.startswith("grad.") or k in ("grad.x1", "grad.x2"):
            continue
        got = named[k[5:]].grad
        got = got[:ROWS] if got.dim() == 2 else got
        if k.endswith("key.bias"):           # analytically zero (softmax is shift invariant): absolute, against a sibling's scale
            gerr[k[5:]] = (got.cpu() - v).abs().max().item() / bias_scale
        else:
            gerr[k[5:]] = norm_err(got, v)
    report("cross_modal_golden", mode, {k: "%.1e" % v for k, v in errs.items()}, {k.replace(".layer.0", ""): "%.1e" % v for k, v in gerr.items()})
    tol = TOL["cross.golden." + mode]
    assert max(errs.values()) < tol["out"], errs
    assert max(gerr.values()) < tol["grad"], gerr


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cross_layer_dropout_replay(mode, report):
    """One layer with all three dropouts on: the masks exported by the library, replayed through the CPU restatement."""
    import mmdti_b200
    from mmdti_b200 import ops, ops_cross
    from tests_util import cross_layer_shapes
    B, L1, L2, D, H, Fd = 3, 19, 23, 256, 8, 512
    p_attn, p_hid, seeds = 0.2, 0.3, (11, 22, 33)
    pre = "layer."
    sd = det_state_dict(cross_layer_shapes(D, Fd, pre), seed=5, std=0.06)
    s1, s2 = det_tensor((B, L1, D), 41, std=1.0), det_tensor((B, L2, D), 42, std=1.0)
    mask2 = torch.ones(B, L2, dtype=torch.bool)
    mask2[1, 9:] = False
    mask2[2, 1:] = False
    up = det_tensor((B, L1, D), 43, std=1.0)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    pc = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    a1, a2 = s1.clone().cuda().requires_grad_(True), s2.clone().cuda().requires_grad_(True)
    P = lambda n: pc[pre + n]
    out = ops_cross.CrossLayerFn.apply(
        a1, a2, mask2.cuda(), P("attention.self.query.weight"), P("attention.self.query.bias"), P("attention.self.key.weight"),
        P("attention.self.key.bias"), P("attention.self.value.weight"), P("attention.self.value.bias"), P("attention.output.dense.weight"),
        P("attention.output.dense.bias"), P("attention.output.LayerNorm.weight"), P("attention.output.LayerNorm.bias"),
        P("intermediate.dense.weight"), P("intermediate.dense.bias"), P("output.dense.weight"), P("output.dense.bias"),
        P("output.LayerNorm.weight"), P("output.LayerNorm.bias"), (H, p_attn, p_hid, seeds, dt, 1e-12))
    (out.float() * up.cuda()).sum().backward()
    keeps = (ops_cross.cross_attn_dropout_mask(B, H, L1, L2, p_attn, seeds[0]).cpu(),
             ops.dropout_mask(B * L1 * D, p_hid, seeds[1]).view(B, L1, D).cpu(), ops.dropout_mask(B * L1 * D, p_hid, seeds[2]).view(B, L1, D).cpu())
    pr = {k: v.clone().double().requires_grad_(True) for k, v in sd.items()}
    r1, r2 = s1.clone().double().requires_grad_(True), s2.clone().double().requires_grad_(True)
    want = restate.cross_layer(r1, r2, mask2, pr, pre, heads=H, eps=1e-12, keeps=keeps, attn_dropout=p_attn, dropout=p_hid)
    (want * up.double()).sum().backward()
    errs = dict(out=rel_err(out, want), ds1=norm_err(a1.grad, r1.grad), ds2=norm_err(a2.grad, r2.grad))
    qb = pr[pre + "attention.self.query.bias"].grad.abs().max().item()
    for k in sd:
        if k.endswith("key.bias"):
            errs[k] = (pc[k].grad.cpu().double() - pr[k].grad).abs().max().item() / qb
        else:
            errs[k] = norm_err(pc[k].grad, pr[k].grad)
    report("cross_layer_dropout_replay", mode, {k.replace(pre, ""): "%.1e" % v for k, v in errs.items()})
    tol = TOL["cross.layer." + mode]
    assert max(errs.values()) < tol, errs


def test_masked_pool_matches_reference_lines():
    from mmdti_b200.models.cross_modal import fuse_and_pool
    B, L1, L2, D = 5, 17, 9, 192
    g = torch.Generator().manual_seed(3)
    for dt in (torch.float32, torch.bfloat16):
        x1 = torch.randn(B, L1, D, generator=g).to(dt).cuda().requires_grad_(True)
        x2 = torch.randn(B, L2, D, generator=g).to(dt).cuda().requires_grad_(True)
        m1 = (torch.rand(B, L1, generator=g) > 0.3).cuda()
        m2 = (torch.rand(B, L2, generator=g) > 0.3).cuda()
        m1[:, 0] = True
        up = torch.randn(B, D, generator=g).cuda()
        got = fuse_and_pool(x1, x2, m1, m2)
        (got * up).sum().backward()
        y1, y2 = x1.detach().double().requires_grad_(True), x2.detach().double().requires_grad_(True)
        want = restate.fuse_pool(y1, y2, m1, m2)
        (want * up.double()).sum().backward()
        assert rel_err(got, want) < 2e-6
        assert rel_err(x1.grad, y1.grad) < (2e-6 if dt == torch.float32 else 5e-3)
        assert rel_err(x2.grad, y2.grad) < (2e-6 if dt == torch.float32 else 5e-3)
