"""GPU parity of K2 (pair-biased attention fwd/bwd, C ABI mmdti_pair_attn_*) vs the oracle."""
import pytest
import torch

from conftest import norm_err, rel_err
from oracle import restate

pytestmark = pytest.mark.gpu

MODES = [("fp32", "fp32"), ("bf16", "bf16"), ("bf16", "fp16"), ("bf16", "fp32")]
TORCH = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}
# tolerances (max-abs error / max-abs reference): fp32 validation mode 1e-5 class; the
# bf16 modes are bounded by bf16 operand rounding (2^-9) of the tensor-core inputs.
from tolerances import TOL as _T

TOL = {"fp32": _T["k2.fp32"], "bf16": _T["k2.bf16"]}


def _problem(B, H, L, act, pair, seed, n_pad):
    g = torch.Generator().manual_seed(seed)
    D = H * 8
    qkv = torch.randn(B * L, 3 * D, generator=g) * 0.7
    bias = torch.randn(B, H, L, L, generator=g)
    pad = torch.zeros(B, L, dtype=torch.bool)
    for b in range(B):
        k = min(n_pad * b, L - 1)
        if k:
            pad[b, L - k:] = True
    bias.masked_fill_(pad[:, None, None, :], float("-inf"))
    d_o = torch.randn(B * L, D, generator=g)
    d_s = torch.randn(B, H, L, L, generator=g) * 0.3
    adt, pdt = TORCH[act], TORCH[pair]
    # round inputs to the storage types so the oracle sees exactly what the kernel sees
    return (qkv.to(adt), bias.to(pdt), d_o.to(adt), d_s.to(pdt), pad)


def _oracle(qkv, bias, d_o, d_s, B, H, L, p, keep):
    D = H * 8
    qkv = qkv.double().cpu().requires_grad_(True)
    bias = bias.double().cpu().requires_grad_(True)
    q, k, v = qkv.view(B, L, 3 * D).chunk(3, dim=-1)
    sp = lambda t: t.reshape(B, L, H, 8).transpose(1, 2)
    o, s = restate.pair_attention(sp(q), sp(k), sp(v), bias, 8 ** -0.5, p, keep)
    o2 = o.transpose(1, 2).reshape(B * L, D)
    fin = torch.isfinite(s)
    loss = (o2 * d_o.double().cpu()).sum() + (torch.where(fin, s, torch.zeros_like(s)) * d_s.double().cpu()).sum()
    loss.backward()
    return o2.detach(), s.detach(), qkv.grad, bias.grad


@pytest.mark.parametrize("act,pair", MODES)
@pytest.mark.parametrize("L", [5, 16, 33, 66, 67, 130, 258])
def test_pair_attn_parity(act, pair, L, report):
    from mmdti_b200 import ops
    B, H = (2, 3) if L > 100 else (3, 4)
    qkv, bias, d_o, d_s, pad = _problem(B, H, L, act, pair, seed=L, n_pad=3)
    dev = "cuda"
    qkv_g = qkv.to(dev).requires_grad_(True)
    bias_g = bias.to(dev).requires_grad_(True)
    Lp = ops.pair_ld(L)
    pair_t = ops.PairPadFn.apply(bias_g.view(B * H, L, L), B, H, L, bias_g.dtype)       # dense -> padded layout
    o, s_pad = ops.pair_attention(qkv_g, pair_t, B, H, L, 8 ** -0.5, 0.0, 0)
    d_s_g = d_s.to(dev).clone()
    d_s_g.masked_fill_(pad.to(dev)[:, None, None, :], 0)
    torch.autograd.backward([o, s_pad], [d_o.to(dev), torch.nn.functional.pad(d_s_g, (0, Lp - L))])
    torch.cuda.synchronize()
    assert torch.isinf(s_pad[..., L:]).all() and (s_pad[..., L:] < 0).all()           # padding columns stay -inf
    s = s_pad[..., :L]
    ro, rs, rdqkv, rdb = _oracle(qkv.float(), bias.float(), d_o.float(), d_s.float(), B, H, L, 0.0, None)
    tol = TOL[act]
    errs = dict(o=rel_err(o.float(), ro), s=rel_err(s.float(), rs), dqkv=rel_err(qkv_g.grad.float(), rdqkv),
                dbias=rel_err(bias_g.grad.float(), rdb), n_dqkv=norm_err(qkv_g.grad.float(), rdqkv),
                n_dbias=norm_err(bias_g.grad.float(), rdb), n_o=norm_err(o.float(), ro))
    report("pair_attn", act, pair, "L=%d" % L, {k: "%.2e" % v for k, v in errs.items()})
    # the stored scores are rounded to the pair dtype
    s_tol = {"fp32": 2e-6, "bf16": tol["s"], "fp16": 1.5e-3}[pair] if act != "fp32" else tol["s"]
    assert errs["o"] < tol["o"], errs
    assert errs["s"] < max(s_tol, tol["s"]), errs
    assert errs["dqkv"] < tol["g"], errs
    assert errs["dbias"] < tol["g"], errs
    # -inf (padded key) pattern is exact and carries no gradient
    assert torch.equal(torch.isinf(s.float().cpu()), pad[:, None, None, :].expand(B, H, L, L))
    assert (bias_g.grad.float().cpu()[pad[:, None, None, :].expand(B, H, L, L)] == 0).all()


@pytest.mark.parametrize("act,pair", [("fp32", "fp32"), ("bf16", "bf16")])
def test_pair_attn_dropout_replay(act, pair, report):
    """Training-mode parity: export the kernel's keep mask and replay it in the oracle."""
    from mmdti_b200 import ops
    B, H, L, p, seed = 2, 4, 66, 0.1, 12345
    qkv, bias, d_o, d_s, pad = _problem(B, H, L, act, pair, seed=7, n_pad=5)
    dev = "cuda"
    qkv_g = qkv.to(dev).requires_grad_(True)
    bias_g = bias.to(dev).requires_grad_(True)
    Lp = ops.pair_ld(L)
    pair_t = ops.PairPadFn.apply(bias_g.view(B * H, L, L), B, H, L, bias_g.dtype)
    o, s = ops.pair_attention(qkv_g, pair_t, B, H, L, 8 ** -0.5, p, seed)
    d_s_g = d_s.to(dev).clone()
    d_s_g.masked_fill_(pad.to(dev)[:, None, None, :], 0)
    torch.autograd.backward([o, s], [d_o.to(dev), torch.nn.functional.pad(d_s_g, (0, Lp - L))])
    keep = ops.attn_dropout_mask(B, H, L, p, seed).cpu()
    rate = keep.float().mean().item()
    thr = round(p * 65536)
    assert abs(rate - (1 - thr / 65536)) < 5e-3, rate
    # the kernel rescales by the exact keep probability 1 - thr/65536
    ro, rs, rdqkv, rdb = _oracle(qkv.float(), bias.float(), d_o.float(), d_s.float(), B, H, L, thr / 65536, keep)
    tol = TOL[act]
    errs = dict(o=rel_err(o.float(), ro), dqkv=rel_err(qkv_g.grad.float(), rdqkv), dbias=rel_err(bias_g.grad.float(), rdb))
    report("pair_attn_dropout", act, pair, {k: "%.2e" % v for k, v in errs.items()}, "keep_rate=%.4f" % rate)
    assert errs["o"] < tol["o"] and errs["dqkv"] < tol["g"] and errs["dbias"] < tol["g"], errs
    # determinism: same seed -> same output; different seed -> different mask
    o2, _ = ops.pair_attention(qkv_g.detach(), pair_t.detach(), B, H, L, 8 ** -0.5, p, seed)
    assert torch.equal(o2, o.detach())
    keep2 = ops.attn_dropout_mask(B, H, L, p, seed + 1).cpu()
    assert (keep2 != keep).float().mean().item() > 0.05


def test_pair_attn_inplace_and_no_dpair(report):
    """pair_out may alias pair_in; d_pair_out may be absent (last layer of MM-DTI)."""
    from mmdti_b200 import ops
    B, H, L = 2, 4, 66
    qkv, bias, d_o, d_s, pad = _problem(B, H, L, "bf16", "bf16", seed=3, n_pad=4)
    dev = "cuda"
    qkv_g = qkv.to(dev).requires_grad_(True)
    bias_g = bias.to(dev).requires_grad_(True)
    pair_t = ops.PairPadFn.apply(bias_g.view(B * H, L, L), B, H, L, bias_g.dtype)
    o, s = ops.pair_attention(qkv_g, pair_t, B, H, L, 8 ** -0.5, 0.0, 0)
    o.backward(d_o.to(dev))
    g1, gb1 = qkv_g.grad.clone(), bias_g.grad.clone()
    ro, rs, rdqkv, rdb = _oracle(qkv.float(), bias.float(), d_o.float(), d_s.float() * 0, B, H, L, 0.0, None)
    assert rel_err(g1.float(), rdqkv) < TOL["bf16"]["g"] and rel_err(gb1.float(), rdb) < TOL["bf16"]["g"]
    with torch.no_grad():
        b2 = pair_t.detach().clone()
        o2, s2 = ops.pair_attention(qkv.to(dev), b2, B, H, L, 8 ** -0.5, 0.0, 0, True)
        assert s2.data_ptr() == b2.data_ptr()
        assert torch.equal(o2, o.detach()) and torch.equal(s2, s.detach())


def test_pair_attn_argument_errors():
    from mmdti_b200 import ops
    from mmdti_b200._lib import MMDTIError
    with pytest.raises(MMDTIError):
        ops.pair_attention(torch.zeros(10, 96), torch.zeros(1, 4, 10, 10), 1, 4, 10, 1.0)      # CPU tensors
    with pytest.raises(MMDTIError):     # qkv of the wrong width
        ops.pair_attention(torch.zeros(10, 90, device="cuda"), torch.zeros(1, 4, 10, 24, device="cuda"), 1, 4, 10, 1.0)
    with pytest.raises(MMDTIError):     # dense instead of padded pair tensor
        ops.pair_attention(torch.zeros(10, 96, device="cuda"), torch.zeros(1, 4, 10, 10, device="cuda"), 1, 4, 10, 1.0)
    with pytest.raises(MMDTIError):     # L beyond the supported maximum
        ops.pair_ld(300)


@pytest.mark.parametrize("L", [5, 66, 130])
def test_pair_pad_unpad_roundtrip(L):
    from mmdti_b200 import ops
    B, H = 2, 3
    x = torch.randn(B * H, L, L)
    for dt in (torch.float32, torch.bfloat16, torch.float16):
        xp = ops.PairPadFn.apply(x.cuda(), B, H, L, dt)
        assert xp.shape == (B, H, L, ops.pair_ld(L)) and ops.pair_ld(L) % 16 == 8
        assert torch.equal(xp[..., :L].cpu().reshape(B * H, L, L), x.to(dt))
        assert torch.isinf(xp[..., L:]).all()
        assert torch.equal(ops.pair_unpad(xp, L, torch.float32).cpu(), x.to(dt).float())


@pytest.mark.parametrize("pdt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("L", [137, 150, 200, 258, 264])
def test_column_split_matches_row_split(L, pdt, monkeypatch):
    """L > 136: the column-split forward (MMDTI_K2_FWD_CS=1) and backward (4 warps per 16-row block, row max / sum / dQ exchanged through shared memory,
    MMDTI_K2_BWD_CS=1) against the one-warp-per-row-block form on the same inputs, dropout on.  Both draw the same mask;
    the only differences are fp32 summation orders (row sums, dQ) ahead of the rounding to the storage types."""
    from mmdti_b200 import _lib, ops
    from mmdti_b200._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64
    B, H, D, p, seed = 2, 64, 512, 0.1, 4321
    Lp = ops.pair_ld(L)
    g = torch.Generator(device="cuda").manual_seed(L)
    qkv = (torch.randn(B * L, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    pair = torch.randn(B, H, L, Lp, device="cuda", generator=g).to(pdt)
    pair[..., L:] = float("-inf")
    pair[1, :, :, L - 3:L] = float("-inf")                                  # masked keys
    d_o = (torch.randn(B * L, D, device="cuda", generator=g) * 0.1).bfloat16()
    dpo = (torch.randn(B, H, L, Lp, device="cuda", generator=g) * 0.01).to(pdt)
    dpo[..., L:] = 0
    dpo[1, :, :, L - 3:L] = 0
    code, pcode = DTYPE_CODE[torch.bfloat16], DTYPE_CODE[pdt]
    fw = []
    for cs in ("1", "0"):                 # the forward has the same two forms (MMDTI_K2_FWD_CS); the row-split outputs feed the backward
        monkeypatch.setenv("MMDTI_K2_FWD_CS", cs)
        pout, o = torch.full_like(pair, 7.0), torch.full((B * L, D), 7.0, device="cuda", dtype=torch.bfloat16)
        call("mmdti_pair_attn_fwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pair, pout, o, i64(D), i32(B), i32(H),
             i32(L), f32(8 ** -0.5), f32(p), u64(seed), i32(code), i32(pcode), stream_ptr())
        torch.cuda.synchronize()
        fw.append((pout, o))
    assert torch.equal(fw[0][0][..., :L], fw[1][0][..., :L])              # P' = scale * QK^T + P: same operations, same bits
    o1, o0 = fw[0][1].float(), fw[1][1].float()
    assert torch.isfinite(o1).all()
    assert (o1 - o0).abs().max() <= 2 ** -6 * o0.abs().max() and (o1 - o0).abs().mean() <= 2e-3 * o0.abs().mean()
    res = []
    for cs in ("0", "1"):
        monkeypatch.setenv("MMDTI_K2_BWD_CS", cs)
        dpi, dqkv = torch.full_like(pair, 7.0), torch.full_like(qkv, 7.0)
        for dpo_ in (dpo, None):                                            # with and without an incoming pair gradient
            call("mmdti_pair_attn_bwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pout, o, d_o, i64(D), dpo_, dpi,
                 dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], i64(3 * D), i32(B), i32(H), i32(L), f32(8 ** -0.5), f32(p),
                 u64(seed), i32(code), i32(pcode), i32(pcode), stream_ptr())
            torch.cuda.synchronize()
            res.append((dpi[..., :L].float().clone(), dqkv.float().clone()))
    for (dp0, dq0), (dp1, dq1) in ((res[0], res[2]), (res[1], res[3])):
        assert torch.isfinite(dp1).all() and torch.isfinite(dq1).all()
        assert (dp0 - dp1).abs().max() <= 2 ** -7 * dp0.abs().max()
        assert (dq0 - dq1).abs().max() <= 2 ** -6 * dq0.abs().max()
        assert (dp0 - dp1).abs().mean() <= 1e-4 * dp0.abs().mean() and (dq0 - dq1).abs().mean() <= 2e-3 * dq0.abs().mean()


@pytest.mark.parametrize("B,L", [(12, 66), (3, 258), (1, 20)])
def test_bwd_dynamic_tile_scheduler_is_bit_identical(B, L, monkeypatch):
    """MMDTI_K2_BWD_DYN hands the last fraction of the (molecule, head) tiles out from a global counter instead of splitting
    all of them into contiguous ranges (the data-parallel bench turns it on).  Which CTA runs a tile does not enter its math:
    every output must be bit-identical for any pool fraction, launch after launch (the counter re-arms itself)."""
    from mmdti_b200 import ops
    from mmdti_b200._lib import DTYPE_CODE, call, f32, i32, i64, stream_ptr, u64
    H, D, p, seed = 64, 512, 0.1, 99
    Lp = ops.pair_ld(L)
    g = torch.Generator(device="cuda").manual_seed(L + B)
    qkv = (torch.randn(B * L, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    pair = torch.randn(B, H, L, Lp, device="cuda", generator=g).bfloat16()
    pair[..., L:] = float("-inf")
    d_o = (torch.randn(B * L, D, device="cuda", generator=g) * 0.1).bfloat16()
    dpo = (torch.randn(B, H, L, Lp, device="cuda", generator=g) * 0.01).bfloat16()
    dpo[..., L:] = 0
    code = DTYPE_CODE[torch.bfloat16]
    pout, o = torch.empty_like(pair), torch.empty((B * L, D), device="cuda", dtype=torch.bfloat16)
    call("mmdti_pair_attn_fwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pair, pout, o, i64(D), i32(B), i32(H), i32(L),
         f32(8 ** -0.5), f32(p), u64(seed), i32(code), i32(code), stream_ptr())
    res = []
    for frac in ("0", "0.3", "1", "0.3"):
        monkeypatch.setenv("MMDTI_K2_BWD_DYN", frac)
        dpi, dqkv = torch.full_like(pair, 7.0), torch.full_like(qkv, 7.0)
        call("mmdti_pair_attn_bwd", qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], i64(3 * D), pout, o, d_o, i64(D), dpo, dpi,
             dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], i64(3 * D), i32(B), i32(H), i32(L), f32(8 ** -0.5), f32(p), u64(seed),
             i32(code), i32(code), i32(code), stream_ptr())
        torch.cuda.synchronize()
        res.append((dpi[..., :L].clone(), dqkv.clone()))
    for dpi, dqkv in res[1:]:
        assert torch.equal(dpi, res[0][0]) and torch.equal(dqkv, res[0][1])
