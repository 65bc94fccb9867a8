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
    if act == "fp32":
        assert e_out < 2e-5
        assert max(e_g.values()) < 2e-4, e_g
    else:
        assert e_out < 2e-2
        assert max(e_g.values()) < 6e-2, e_g


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
