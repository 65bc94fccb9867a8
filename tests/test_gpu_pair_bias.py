"""GPU parity of K1 (pair bias fwd/bwd, mask fill, pair outputs) vs golden fixtures + oracle."""
import pytest
import torch

from conftest import load_golden, norm_err, rel_err
from oracle import restate

pytestmark = pytest.mark.gpu


def _mods(g, dev):
    from mmdti_b200.models.encoder import GaussianLayer, NonLinearHead
    gbf, proj = GaussianLayer(128, 961), NonLinearHead(128, 64, "gelu")
    gbf.load_state_dict({k[len("w.gbf."):]: v for k, v in g.items() if k.startswith("w.gbf.")})
    proj.load_state_dict({k[len("w.gbf_proj."):]: v for k, v in g.items() if k.startswith("w.gbf_proj.")})
    return gbf.to(dev), proj.to(dev)


@pytest.mark.parametrize("tag", ["init", "pre"])
@pytest.mark.parametrize("act,pair", [("fp32", "fp32"), ("bf16", "bf16"), ("bf16", "fp32")])
def test_pair_bias_golden(tag, act, pair, report):
    """Module-level drop-in: the reference call sequence gbf -> gbf_proj -> permute ->
    contiguous (models/mm_model.py:553-556) against the fixture made by the reference."""
    import mmdti_b200
    g = load_golden("pair_bias_" + tag)
    dev = "cuda"
    with mmdti_b200.precision(act=act, pair=pair):
        gbf, proj = _mods(g, dev)
        out = proj(gbf(g["in.dist"].to(dev), g["in.edge_type"].to(dev)))
        out = out.permute(0, 3, 1, 2).contiguous()
        assert out.is_contiguous() and out.shape == g["out.bias"].shape
        (out.float() * g["in.upstream"].to(dev)).sum().backward()
    e_out = rel_err(out.float(), g["out.bias"])
    grads = {"gbf." + k: v.grad for k, v in gbf.named_parameters()}
    grads.update({"gbf_proj." + k: v.grad for k, v in proj.named_parameters()})
    e_g = {k: rel_err(grads[k].float(), g["grad." + k]) for k in grads}
    report("pair_bias", tag, act, pair, "out=%.2e" % e_out, {k: "%.1e" % v for k, v in e_g.items()})
    from tolerances import TOL
    t = TOL["k1." + act]
    assert e_out < t["out"], e_out
    assert max(e_g.values()) < t["grad"], e_g


@pytest.mark.parametrize("B,n_atoms", [(128, 64), (8, 256)])
@pytest.mark.parametrize("tag", ["init", "pre"])
def test_pair_bias_bench_sizes(B, n_atoms, tag, report):
    """K1 forward + backward at the sizes bench.py runs (config 2: 128 x 66^2 = 557 568 pairs; config 4 shape:
    8 x 258^2 = 532 512 pairs), so that the persistent multi-tile loops of csrc/pair_bias.cu (64 pairs per tile,
    grid <= 444 => ~20 tiles per CTA) and csrc/pair_bias_bwd.cu (128 pairs per tile, grid <= 148 => ~29 tiles per CTA,
    dW1 / dW2 register accumulators carried across all of them) iterate many times under a checker.
    Truth = oracle/restate.pair_bias in float64 (``wide=True``), evaluated molecule-chunk by molecule-chunk."""
    import mmdti_b200
    from mmdti_b200 import ops
    from mmdti_b200.data import synthetic_molecules
    from tolerances import TOL
    g = load_golden("pair_bias_" + tag)
    dev = "cuda"
    tokens, dist, et, _ = synthetic_molecules(B, n_atoms, seed=77, ragged=True)
    L = tokens.shape[1]
    pad = tokens.eq(0)
    H = 64
    up = torch.randn(B, H, L, L, generator=torch.Generator().manual_seed(3))
    p64 = {k[2:]: v.double().requires_grad_(True) for k, v in g.items() if k.startswith("w.")}
    want = torch.empty(B, H, L, L)
    step = max(1, 40000 // (L * L))
    for b0 in range(0, B, step):
        sl = slice(b0, min(B, b0 + step))
        o = restate.pair_bias(dist[sl].double(), et[sl], p64, wide=True).view(-1, H, L, L)
        want[sl] = o.detach().float()
        (o * up[sl].double().masked_fill(pad[sl][:, None, None, :], 0.0)).sum().backward()
    want.masked_fill_(pad[:, None, None, :], float("-inf"))
    for act, pair in (("fp32", "fp32"), ("bf16", "bf16")):
        with mmdti_b200.precision(act=act, pair=pair):
            gbf, proj = _mods(g, dev)
            pj = proj
            out = ops.pair_bias(dist.to(dev), et.to(dev), gbf.means.weight, gbf.stds.weight, gbf.mul.weight, gbf.bias.weight,
                                pj.linear1.weight, pj.linear1.bias, pj.linear2.weight, pj.linear2.bias, key_pad=pad.to(dev))
            assert out.shape == (B, H, L, ops.pair_ld(L))
            # padding columns of the layout hold -inf and receive no gradient
            upp = torch.zeros_like(out, dtype=torch.float32)
            upp[..., :L] = up.to(dev)
            (out.float().masked_fill(torch.isinf(out), 0.0) * upp).sum().backward()
        got = out[..., :L].float().cpu()
        e_out = rel_err(got, want)
        grads = {"gbf." + k: v.grad for k, v in gbf.named_parameters()}
        grads.update({"gbf_proj." + k: v.grad for k, v in proj.named_parameters()})
        e_g = {k: norm_err(grads[k].float().cpu().view(-1), p64[k].grad.float().view(-1)) for k in grads}
        report("pair_bias_bench_size", "B=%d L=%d" % (B, L), tag, act, "out=%.2e" % e_out, {k: "%.1e" % v for k, v in e_g.items()})
        t = TOL["k1." + act]
        assert e_out < t["out"], e_out
        assert max(e_g.values()) < t["grad"], e_g


def test_pair_bias_keypad_fused_and_mask_fill(report):
    """-inf merge of the key-padding mask: fused in K1 == the in-place kernel == the oracle's
    masked_fill_ (models/transformers.py:122-132), bit-exact pattern."""
    import mmdti_b200
    from mmdti_b200 import ops
    from mmdti_b200.data import synthetic_molecules
    g = load_golden("pair_bias_pre")
    dev = "cuda"
    tokens, dist, et, _ = synthetic_molecules(4, 13, seed=5, ragged=True)
    pad = tokens.eq(0)
    assert pad.any()
    for act, pair in (("fp32", "fp32"), ("bf16", "bf16"), ("bf16", "fp16")):
        with mmdti_b200.precision(act=act, pair=pair):
            gbf, proj = _mods(g, dev)
            fused = proj(gbf(dist.to(dev), et.to(dev)), key_pad=pad.to(dev)).permute(0, 3, 1, 2).contiguous()
            plain = proj(gbf(dist.to(dev), et.to(dev))).permute(0, 3, 1, 2).contiguous()
            filled = ops.pair_mask_fill_(plain.clone().view(-1, 15, 15), pad.to(dev)).view_as(plain)
        want = pad[:, None, None, :].expand(4, 64, 15, 15)
        assert torch.equal(torch.isinf(fused.float().cpu()), want)
        assert torch.equal(fused.cpu(), filled.cpu())
        # untouched entries are bit-identical to the unmasked tensor
        assert torch.equal(filled.cpu()[~want], plain.cpu()[~want])
    p = {k[2:]: v for k, v in g.items() if k.startswith("w.")}
    ref = restate.merge_key_padding(restate.pair_bias(dist, et, p), pad, 64).view(4, 64, 15, 15)
    assert rel_err(fused.float(), ref) < 2e-2


def test_pair_outputs(report):
    from mmdti_b200 import ops
    B, H, L = 2, 8, 11
    g = torch.Generator().manual_seed(0)
    first = torch.randn(B, H, L, L, generator=g)
    pad = torch.zeros(B, L, dtype=torch.bool)
    pad[1, -3:] = True
    first.masked_fill_(pad[:, None, None, :], float("-inf"))
    last = first + torch.randn(B, H, L, L, generator=g)
    fp = ops.PairPadFn.apply(first.cuda().view(B * H, L, L), B, H, L, torch.float32)
    lp = ops.PairPadFn.apply(last.cuda().view(B * H, L, L), B, H, L, torch.float32)
    pair, delta = ops.PairOutputsFn.apply(fp, lp, B, H, L)
    want_delta = (last - first).masked_fill(pad[:, None, None, :], 0).permute(0, 2, 3, 1)
    assert torch.equal(pair.cpu(), last.permute(0, 2, 3, 1))
    assert torch.equal(delta.cpu(), want_delta)
