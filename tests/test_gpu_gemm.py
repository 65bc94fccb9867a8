"""GPU parity of the tcgen05 projection GEMMs with fused epilogues (csrc/gemm_tc.cu) against plain PyTorch fp32 math on
the same bf16 operands.  The fused epilogues restate Uni-Core's pre-LN TransformerEncoderLayer (SURVEY.md Appendix A;
reference call sites models/transformers.py:82-91,136-139): bias, exact-erf GELU, dropout + residual + LayerNorm and
their backward passes.  Dropout masks are replayed from the library's own debug export (ops.dropout_mask), so the
comparison is deterministic.  Operand rounding is identical on both sides (bf16 inputs); the tolerance covers the
accumulation order (fp32) and the bf16 rounding of the stored outputs (2^-9 relative)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(264, 512, 1536, 2048), (8448, 512, 1536, 2048), (200, 64, 192, 128), (1000, 256, 768, 1024)]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale)


def _close(a, b, tol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)
    nerr = ((a - b).norm() / b.norm().clamp_min(1e-20)).item()
    assert err < tol and nerr < tol, (what, err, nerr)
    return err


@pytest.mark.parametrize("M,D,D3,Fd", SHAPES)
def test_forward_gemms(M, D, D3, Fd, report):
    from mmdti_b200 import ops, ops_gemm
    x = _rand((M, D), 1).bfloat16()
    w_in, b_in = _rand((D3, D), 2, 0.05).bfloat16(), _rand((D3,), 3, 0.1).bfloat16()
    # in_proj: bias epilogue
    y = ops_gemm.gemm_bias(x, w_in, b_in)
    want = x.float() @ w_in.float().t() + b_in.float()
    e1 = _close(y, want, 6e-3, "gemm_bias")
    # the output may be written into a strided view (q|k|v columns of a wider buffer)
    buf = torch.zeros((M, D3 + 64), device="cuda", dtype=torch.bfloat16)
    ops_gemm.gemm_bias(x, w_in, b_in, out=buf[:, :D3])
    assert torch.equal(buf[:, :D3], y) and buf[:, D3:].abs().sum() == 0
    # fc1: bias + GELU, both outputs
    w1, b1 = _rand((Fd, D), 4, 0.05).bfloat16(), _rand((Fd,), 5, 0.1).bfloat16()
    z, u = ops_gemm.gemm_bias_gelu(x, w1, b1)
    zw = x.float() @ w1.float().t() + b1.float()
    e2 = _close(z, zw, 6e-3, "fc1.z")
    e3 = _close(u, F.gelu(z.float()), 6e-3, "fc1.u (gelu of the stored z)")
    # store_grad: the first buffer receives gelu'(z) (what the fused backward multiplies by), u = gelu of the fp32 z
    gp, u2 = ops_gemm.gemm_bias_gelu(x, w1, b1, store_grad=True)
    zr = zw.clone().requires_grad_(True)
    F.gelu(zr).sum().backward()
    e4 = _close(gp, zr.grad, 6e-3, "fc1.gelu'")
    e5 = _close(u2, F.gelu(zw), 6e-3, "fc1.u (gelu of the fp32 z)")
    report("gemm_fwd", (M, D, D3, Fd), "bias=%.1e z=%.1e u=%.1e gelu'=%.1e u2=%.1e" % (e1, e2, e3, e4, e5))


@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("M,D,D3,Fd", SHAPES)
def test_dropres_layernorm_epilogue(M, D, D3, Fd, p, report):
    """out_proj / fc2: xo = res + dropout(x W^T + b), y = LayerNorm(xo), statistics saved; K = D and K = Fd"""
    from mmdti_b200 import ops, ops_gemm
    for K, seed in ((D, 11), (Fd, 12)):
        x = _rand((M, K), seed).bfloat16()
        w, b = _rand((D, K), seed + 1, 0.05).bfloat16(), _rand((D,), seed + 2, 0.1).bfloat16()
        res = _rand((M, D), seed + 3)
        ln_w, ln_b = 1 + 0.1 * _rand((D,), seed + 4), 0.1 * _rand((D,), seed + 5)
        xo, y, st = ops_gemm.gemm_dropres_ln(x, w, b, res, ln_w, ln_b, p, 777 + seed)
        a = x.float() @ w.float().t() + b.float()
        if p > 0:
            keep = ops.dropout_mask(M * D, p, 777 + seed).view(M, D)
            a = a * keep / (1 - p)
        xo_w = res + a
        e1 = _close(xo, xo_w, 2e-5 if K == D else 1e-4, "xo")
        mu, var = xo_w.mean(1), xo_w.var(1, unbiased=False)
        y_w = (xo_w - mu[:, None]) * torch.rsqrt(var + 1e-5)[:, None] * ln_w + ln_b
        e2 = _close(y, y_w, 6e-3, "ln out")
        e3 = _close(st[0], mu, 1e-4, "mean")
        e4 = _close(st[1], torch.rsqrt(var + 1e-5), 1e-4, "rstd")
        # without LayerNorm (the last layer's fc2)
        xo2, y2, st2 = ops_gemm.gemm_dropres_ln(x, w, b, res, None, None, p, 777 + seed)
        assert y2 is None and torch.equal(xo2, xo)
        report("gemm_dropres_ln", (M, D, K), "p=%.1f xo=%.1e y=%.1e mean=%.1e rstd=%.1e" % (p, e1, e2, e3, e4))


@pytest.mark.parametrize("M,D,D3,Fd", SHAPES)
def test_backward_gemms(M, D, D3, Fd, report):
    from mmdti_b200 import ops, ops_gemm
    # plain dgrad (out_proj): d_o = da W_out
    da = _rand((M, D), 21, 0.1).bfloat16()
    w_out = _rand((D, D), 22, 0.05).bfloat16()
    e1 = _close(ops_gemm.gemm_dgrad(da, w_out), da.float() @ w_out.float(), 6e-3, "dgrad")
    # dgrad fc2 + GELU backward + bias column sums
    df = _rand((M, D), 23, 0.1).bfloat16()
    w2 = _rand((D, Fd), 24, 0.05).bfloat16()
    z = _rand((M, Fd), 25).bfloat16()
    dbias = torch.zeros(Fd, device="cuda")
    dz = ops_gemm.gemm_dgrad_gelu(df, w2, z, dbias)
    zf = z.float().requires_grad_(True)
    F.gelu(zf).backward(df.float() @ w2.float())
    e2 = _close(dz, zf.grad, 6e-3, "dgrad_gelu")
    e3 = _close(dbias, dz.float().sum(0), 2e-4, "dgrad_gelu.dbias == colsum of the stored dz")
    # z_is_grad: the buffer already holds gelu'(z)
    gp = zf.grad.new_tensor(0)          # placeholder
    zg = z.float().clone().requires_grad_(True)
    F.gelu(zg).sum().backward()
    dbias2 = torch.zeros(Fd, device="cuda")
    dz2 = ops_gemm.gemm_dgrad_gelu(df, w2, zg.grad.bfloat16(), dbias2, z_is_grad=True)
    _close(dz2, (df.float() @ w2.float()) * zg.grad.bfloat16().float(), 6e-3, "dgrad_gelu with the stored derivative")
    _close(dbias2, dz2.float().sum(0), 2e-4, "dgrad_gelu(z_is_grad).dbias")
    # wgrad: fp32, split over the tokens; plain and accumulating
    h = _rand((M, D), 26).bfloat16()
    dw = ops_gemm.gemm_wgrad(dz, h)
    want = dz.float().t() @ h.float()
    e4 = _close(dw, want, 2e-4, "wgrad")
    ops_gemm.gemm_wgrad(dz, h, out=dw, accumulate=True)
    e5 = _close(dw, 2 * want, 2e-4, "wgrad accumulate")
    # strided operands (column slices of qkv-like buffers)
    big = _rand((M, D3), 27, 0.1).bfloat16()
    dwq = ops_gemm.gemm_wgrad(big[:, D:2 * D], h)
    e6 = _close(dwq, big[:, D:2 * D].float().t() @ h.float(), 2e-4, "wgrad strided")
    report("gemm_bwd", (M, D, D3, Fd), "dgrad=%.1e dgelu=%.1e dbias=%.1e wgrad=%.1e acc=%.1e strided=%.1e" % (e1, e2, e3, e4, e5, e6))


@pytest.mark.parametrize("p", [0.0, 0.1])
@pytest.mark.parametrize("M,D,D3,Fd", SHAPES)
def test_layernorm_backward_epilogue(M, D, D3, Fd, p, report):
    """fc1 / in_proj dgrad: dh = dY W, dx = dx_add + LN'(dh), da = dropout'(dx), dw / db / dbias column sums"""
    from mmdti_b200 import ops, ops_gemm
    for N, seed in ((Fd, 31), (D3, 32)):
        dy = _rand((M, N), seed, 0.1).bfloat16()
        w = _rand((N, D), seed + 1, 0.05).bfloat16()
        x = _rand((M, D), seed + 2) + 0.3
        ln_w = 1 + 0.1 * _rand((D,), seed + 3)
        dx_add = _rand((M, D), seed + 4, 0.1)
        mu, var = x.mean(1), x.var(1, unbiased=False)
        stats = torch.stack([mu, torch.rsqrt(var + 1e-5)]).contiguous()
        dw, db, dbias = (torch.zeros(D, device="cuda") for _ in range(3))
        for add in (dx_add, None):
            dw.zero_(), db.zero_(), dbias.zero_()
            dx, da = ops_gemm.gemm_dgrad_lnbwd(dy, w, x, stats, ln_w, add, dw, db, dbias, p, 999 + seed)
            xr = x.clone().requires_grad_(True)
            wr = ln_w.clone().requires_grad_(True)
            br = torch.zeros(D, device="cuda", requires_grad=True)
            dh = dy.float() @ w.float()
            F.layer_norm(xr, (D,), wr, br, 1e-5).backward(dh)
            dx_w = xr.grad + (add if add is not None else 0)
            e1 = _close(dx, dx_w, 1e-4, "dx")
            da_w = dx_w
            if p > 0:
                da_w = dx_w * ops.dropout_mask(M * D, p, 999 + seed).view(M, D) / (1 - p)
            e2 = _close(da, da_w, 6e-3, "da")
            e3 = _close(dw, wr.grad, 3e-4, "d ln_w")
            e4 = _close(db, br.grad, 3e-4, "d ln_b")
            e5 = _close(dbias, da.float().sum(0), 3e-4, "dbias == colsum of the stored da")
        report("gemm_lnbwd", (M, D, N), "p=%.1f dx=%.1e da=%.1e dw=%.1e db=%.1e dbias=%.1e" % (p, e1, e2, e3, e4, e5))
