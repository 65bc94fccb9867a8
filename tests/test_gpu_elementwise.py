"""GPU parity of the fused elementwise kernels and of the fused encoder layer (training mode,
dropout masks exported from the kernels and replayed in the oracle)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import norm_err, rel_err
from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [32, 64, 512, 1024])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(D, dt):
    from mmdti_b200 import ops
    from mmdti_b200._lib import DTYPE_CODE, call, i32, stream_ptr
    rows = 77
    g = torch.Generator().manual_seed(D)
    x = torch.randn(rows, D, generator=g) * 2 + 0.5
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    dy = torch.randn(rows, D, generator=g).to(dt)
    add = torch.randn(rows, D, generator=g)
    y, st = ops.layernorm_fwd(x.cuda(), w.cuda(), b.cuda(), dt)
    xr = x.double().requires_grad_(True)
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), wr, br, 1e-5)
    (yr * dy.double()).sum().backward()
    assert rel_err(y.float(), yr) < (2e-6 if dt == torch.float32 else 6e-3)
    dx = torch.empty(rows, D, device="cuda")
    dw, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    call("mmdti_layernorm_bwd", dy.cuda(), x.cuda(), w.cuda(), st[0], st[1], add.cuda(), dx, dw, db, i32(rows), i32(D),
         i32(DTYPE_CODE[dt]), stream_ptr())
    assert rel_err(dx, xr.grad + add.double()) < 1e-5
    assert rel_err(dw, wr.grad) < 1e-5 and rel_err(db, br.grad) < 1e-5


def test_dropout_residual_gelu_colsum():
    from mmdti_b200 import ops
    from mmdti_b200._lib import call, f32, i32, i64, stream_ptr, u64
    rows, C, p, seed = 50, 512, 0.1, 99
    g = torch.Generator().manual_seed(1)
    res, a = torch.randn(rows, C, generator=g), torch.randn(rows, C, generator=g)
    keep = ops.dropout_mask(rows * C, p, seed).view(rows, C).cpu()
    thr = round(p * 65536)
    scale = 65536.0 / (65536 - thr)
    assert abs(keep.float().mean().item() - (1 - thr / 65536)) < 1e-2
    for dt, code, tol in ((torch.float32, 0, 1e-6), (torch.bfloat16, 1, 1e-2)):
        out = torch.empty(rows, C, device="cuda")
        call("mmdti_dropout_residual_fwd", res.cuda(), a.to(dt).cuda(), out, i64(rows * C), f32(p), u64(seed), i32(code), stream_ptr())
        want = res + a.to(dt).float() * keep * scale
        assert rel_err(out, want) < 1e-6
        dx = torch.randn(rows, C, generator=g)
        da = torch.empty(rows, C, device="cuda", dtype=dt)
        dbias = torch.zeros(C, device="cuda")
        call("mmdti_dropout_bwd", dx.cuda(), da, dbias, i32(rows), i32(C), f32(p), u64(seed), i32(code), stream_ptr())
        assert rel_err(da.float(), dx * keep * scale) < tol
        assert rel_err(dbias, da.float().sum(0)) < 1e-5
        # GELU (exact erf) fwd/bwd + fused bias gradient
        z = (torch.randn(rows, 2048, generator=g) * 2).to(dt)
        u = torch.empty_like(z, device="cuda")
        call("mmdti_gelu_fwd", z.cuda(), u, i64(z.numel()), i32(code), stream_ptr())
        assert rel_err(u.float(), F.gelu(z.double())) < tol
        du = torch.randn(rows, 2048, generator=g).to(dt)
        dz = torch.empty_like(u)
        dbz = torch.zeros(2048, device="cuda")
        call("mmdti_gelu_bwd", du.cuda(), z.cuda(), dz, dbz, i32(rows), i32(2048), i32(code), stream_ptr())
        zr = z.double().requires_grad_(True)
        (F.gelu(zr) * du.double()).sum().backward()
        assert rel_err(dz.float(), zr.grad) < tol
        assert rel_err(dbz, dz.float().sum(0)) < 1e-5
        cs = torch.zeros(1536, device="cuda")
        xx = torch.randn(rows, 1536, generator=g).to(dt)
        call("mmdti_colsum", xx.cuda(), cs, i32(rows), i32(1536), i32(code), stream_ptr())
        assert rel_err(cs, xx.float().sum(0)) < 1e-5


@pytest.mark.parametrize("act,pair", [("fp32", "fp32"), ("bf16", "bf16")])
def test_fused_layer_training_mode_replay(act, pair, report):
    """One encoder layer in TRAINING mode (all three dropouts on): the kernels' keep masks are
    exported and replayed in the oracle layer, so outputs and gradients must agree."""
    import mmdti_b200
    from mmdti_b200 import ops
    from mmdti_b200.models.unicore_compat import TransformerEncoderLayer
    B, L, H, D, Fd = 3, 21, 8, 64, 128
    p_attn = p_drop = 0.1
    torch.manual_seed(5)
    layer = TransformerEncoderLayer(D, Fd, H, dropout=p_drop, attention_dropout=p_attn).cuda().train()
    with torch.no_grad():
        for n_, p_ in layer.named_parameters():
            p_.normal_(0, 0.2) if "weight" in n_ and "layer_norm" not in n_ else p_.add_(0.1 * torch.randn_like(p_))
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, L, D, generator=g)
    bias = torch.randn(B * H, L, L, generator=g)
    pad = torch.zeros(B, L, dtype=torch.bool)
    pad[1, -4:] = True
    bias.view(B, H, L, L).masked_fill_(pad[:, None, None, :], float("-inf"))
    ux, up = torch.randn(B, L, D, generator=g), torch.randn(B, H, L, L, generator=g) * 0.2
    up.masked_fill_(pad[:, None, None, :], 0)
    with mmdti_b200.precision(act=act, pair=pair):
        pdt = mmdti_b200.config.pair_dtype()
        torch.manual_seed(11)
        ops.reset_seed_counter()
        seeds = [ops.next_seed() for _ in range(3)]
        ops.reset_seed_counter()
        xg = x.cuda().requires_grad_(True)
        bg = bias.cuda().requires_grad_(True)
        pair_t = ops.PairPadFn.apply(bg, B, H, L, pdt)
        y, s, _ = layer(xg, attn_bias=pair_t, return_attn=True)
        Lp = ops.pair_ld(L)
        torch.autograd.backward([y, s], [ux.cuda(), F.pad(up, (0, Lp - L)).to(pdt).cuda()])
    thr = round(0.1 * 65536)
    pe = thr / 65536
    keeps = {"attn": ops.attn_dropout_mask(B, H, L, p_attn, seeds[0]).cpu(),
             "res1": ops.dropout_mask(B * L * D, p_drop, seeds[1]).view(B, L, D).cpu(),
             "res2": ops.dropout_mask(B * L * D, p_drop, seeds[2]).view(B, L, D).cpu()}
    prm = {k: v.detach().cpu().double().requires_grad_(True) for k, v in layer.state_dict().items()}
    xr = x.double().requires_grad_(True)
    br = bias.to(pdt).double().requires_grad_(True)
    yr, sr = restate.encoder_layer(xr, br, prm, "", H, pe, pe, keeps)
    fin = torch.isfinite(sr)
    ((yr * ux.double()).sum() + (torch.where(fin, sr, torch.zeros_like(sr)).view(B, H, L, L) * up.double()).sum()).backward()
    errs = dict(y=rel_err(y, yr), s=rel_err(s[..., :L].float().reshape(B * H, L, L), sr), dx=rel_err(xg.grad, xr.grad),
                dbias=rel_err(bg.grad, br.grad))
    named = dict(layer.named_parameters())
    for k in ("self_attn.in_proj.weight", "self_attn.in_proj.bias", "self_attn.out_proj.weight", "fc1.weight", "fc1.bias",
              "fc2.weight", "fc2.bias", "self_attn_layer_norm.weight", "final_layer_norm.bias", "self_attn.out_proj.bias"):
        errs["d_" + k] = rel_err(named[k].grad, prm[k].grad)
    report("fused_layer_train", act, pair, {k: "%.1e" % v for k, v in errs.items()})
    assert max(errs.values()) < (5e-5 if act == "fp32" else 4e-2), errs


@pytest.mark.parametrize("D", [64, 512])
@pytest.mark.parametrize("dt,code", [(torch.float32, 0), (torch.bfloat16, 1)])
def test_fused_dropres_layernorm_pairs(D, dt, code):
    """dropout+residual+LayerNorm (fwd) and LayerNorm-bwd+dropout-bwd equal the two-kernel sequences
    they replace (same masks, exported by mmdti_dropout_mask)."""
    from mmdti_b200 import ops
    from mmdti_b200._lib import call, f32, i32, stream_ptr, u64
    rows, p, seed = 91, 0.1, 4242
    g = torch.Generator().manual_seed(D + code)
    res, a = torch.randn(rows, D, generator=g), torch.randn(rows, D, generator=g).to(dt)
    w, b = torch.randn(D, generator=g), torch.randn(D, generator=g)
    keep = ops.dropout_mask(rows * D, p, seed).view(rows, D).cpu()
    thr = round(p * 65536)
    scale = 65536.0 / (65536 - thr)
    xo = torch.empty(rows, D, device="cuda")
    y = torch.empty(rows, D, device="cuda", dtype=dt)
    st = torch.empty(2, rows, device="cuda")
    call("mmdti_dropres_layernorm_fwd", res.cuda(), a.cuda(), xo, w.cuda(), b.cuda(), y, st[0], st[1], i32(rows), i32(D), f32(1e-5),
         f32(p), u64(seed), i32(code), i32(code), stream_ptr())
    want_x = res + a.float() * keep * scale
    assert rel_err(xo, want_x) < 1e-6
    want_y = F.layer_norm(want_x.double(), (D,), w.double(), b.double(), 1e-5)
    assert rel_err(y.float(), want_y) < (2e-6 if dt == torch.float32 else 6e-3)
    # backward pair, against mmdti_layernorm_bwd followed by mmdti_dropout_bwd
    dy = torch.randn(rows, D, generator=g).to(dt)
    add = torch.randn(rows, D, generator=g)
    dx_ref = torch.empty(rows, D, device="cuda")
    dw_ref, db_ref = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    call("mmdti_layernorm_bwd", dy.cuda(), xo, w.cuda(), st[0], st[1], add.cuda(), dx_ref, dw_ref, db_ref, i32(rows), i32(D),
         i32(code), stream_ptr())
    da_ref = torch.empty(rows, D, device="cuda", dtype=dt)
    dbias_ref = torch.zeros(D, device="cuda")
    call("mmdti_dropout_bwd", dx_ref, da_ref, dbias_ref, i32(rows), i32(D), f32(p), u64(seed), i32(code), stream_ptr())
    dx = torch.empty(rows, D, device="cuda")
    dw, db, dbias = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    da = torch.empty(rows, D, device="cuda", dtype=dt)
    call("mmdti_layernorm_bwd_dropout", dy.cuda(), xo, w.cuda(), st[0], st[1], add.cuda(), dx, dw, db, da, dbias, i32(rows), i32(D),
         f32(p), u64(seed), i32(code), stream_ptr())
    assert rel_err(dx, dx_ref) < 1e-6 and rel_err(dw, dw_ref) < 1e-5 and rel_err(db, db_ref) < 1e-5
    assert torch.equal(da, da_ref)
    assert rel_err(dbias, dbias_ref) < 1e-5
