"""GPU parity of the ChemBERTa (HF RoBERTa) encoder on the mmdti kernels against the fixture made by
Hugging Face's own ``RobertaModel`` (oracle/make_golden.py:gold_chemberta; the reference loads it at models/mm_model.py:475 and
calls it at :562).  fp32 validation mode at the 1e-5 class, bf16 per tests/tolerances.py."""
import pytest
import torch

from conftest import load_golden, norm_err, rel_err
from oracle.detw import det_state_dict
from tolerances import TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_chemberta_golden(mode, report):
    import mmdti_b200
    from transformers import RobertaConfig
    from mmdti_b200.models.encoder import ChembertaEncoder
    g = load_golden("chemberta")
    H, D, Fd, nl, V, P, seed, ROWS = [int(v) for v in g["cfg"]]
    cfg = RobertaConfig(vocab_size=V, hidden_size=D, num_hidden_layers=nl, num_attention_heads=H, intermediate_size=Fd,
                        max_position_embeddings=P, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pad_token_id=1)
    enc = ChembertaEncoder(cfg)
    assert type(enc.bert).__module__.startswith("mmdti_b200")
    sd = det_state_dict({k: tuple(v.shape) for k, v in enc.bert.state_dict().items()}, seed=seed, std=0.05)
    enc.bert.load_state_dict(sd, strict=True)
    enc = enc.cuda().train()                      # dropout rates are 0: train mode exercises the seeded paths
    ids, am = g["in.ids"].cuda(), g["in.mask"].cuda()
    with mmdti_b200.precision(act=mode, pair=mode):
        out = enc(ids, am)
        (out.float() * g["in.up"].cuda()).sum().backward()
    valid = am.bool()
    errs = {"hidden": rel_err(out.float()[valid], g["out.hidden"].cuda()[valid])}
    named = dict(enc.bert.named_parameters())
    for k, v in g.items():
        if k.startswith("grad."):
            got = named[k[5:]].grad
            got = got[:ROWS] if (got.dim() == 2 and "embeddings" not in k) else got
            errs[k[5:]] = norm_err(got, v)
    report("chemberta_golden", mode, {k: "%.1e" % v for k, v in errs.items()})
    tol = TOL["chemberta." + mode]
    assert errs.pop("hidden") < tol["out"]
    assert max(errs.values()) < tol["grad"], errs
