"""SURVEY.md §4 test (iv): the reference's OWN ``MM_Model`` (models/mm_model.py:408-618) built twice -- with the reference's
classes (Uni-Core restated in oracle/shims) and with the mmdti_b200 drop-in modules patched into the reference's module
namespace (``TransformerEncoderWithPair``, ``GaussianLayer``, ``NonLinearHead``, ``InfoNCE``, ``CT_*``) -- same weights
(``state_dict`` copied with strict=True: the names are the contract of SURVEY.md §8b), same batch, dropout 0; compares
``logits, ct_loss, rnc_loss`` of ``MM_Model.forward(..., return_infonce_loss=True, return_ct_loss=True)`` and gradients.

Needs a copy of the reference tree: /root/reference in the build container, or baseline/_ref (made by
``scripts/install_reference.sh`` / ``__graft_entry__.build()``; git-ignored, travels with the snapshot to the GPU box).
Skipped when neither is present -- the kernel-level parity tests against the committed fixtures do not depend on it."""
import pytest
import torch

from conftest import norm_err, rel_err
from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not ref_loader.available(), reason="no copy of the reference tree on this machine")
@pytest.mark.parametrize("task", ["regression", "classification"])
def test_reference_mm_model_with_dropin_modules(task, report):
    import mmdti_b200
    import mm_model_harness as h
    ref = ref_loader.load()
    dev = "cuda"
    m_ref = h.build(ref, dropin=False, task=task).to(dev).train()
    m_new = h.build(ref, dropin=True, task=task).to(dev).train()
    m_new.load_state_dict(m_ref.state_dict(), strict=True)          # identical state_dict names: the drop-in contract
    assert type(m_new.cross_modal_module).__module__.startswith("mmdti_b200") and type(m_new.bert).__module__.startswith("mmdti_b200")
    assert type(m_new.encoder).__module__.startswith("mmdti_b200") and type(m_ref.encoder).__module__ == "models.transformers"
    inp, y, w = h.batch()
    if task == "classification":
        y = (y > 0).float()
    inp = {k: v.to(dev) for k, v in inp.items()}
    y, w = y.to(dev), w.to(dev)
    kw = dict(weights=w, return_infonce_loss=True, return_ct_loss=True, net_target=y, use_weight=(task == "regression"), epoch=1)

    def run(model):
        model.zero_grad()
        logits, ct_loss, rnc_loss = model(**inp, **kw)
        (logits.float().pow(2).mean() + 0.1 * ct_loss + 0.1 * rnc_loss).backward()
        return logits.detach().float(), ct_loss.detach().float(), rnc_loss.detach().float()

    want = run(m_ref)
    names = ["embed_tokens.weight", "gbf.means.weight", "gbf_proj.linear1.weight", "encoder.layers.0.self_attn.in_proj.weight",
             "encoder.layers.1.fc2.weight", "infonce.info_proj_query.0.weight", "cross_modal_module.text_attention.layer.0.attention.self.query.weight",
             "cross_modal_module.graph_attention.layer.0.output.dense.weight", "classification_head.dense.weight",
             "bert.embeddings.word_embeddings.weight", "bert.encoder.layer.0.attention.self.key.weight",
             "bert.encoder.layer.1.output.dense.weight"]
    g_ref = {k: dict(m_ref.named_parameters())[k].grad.detach().clone() for k in names}
    for act, pair, ltol, gtol in (("fp32", "fp32", 2e-5, 2e-4), ("bf16", "bf16", 2e-2, 6e-2)):
        with mmdti_b200.precision(act=act, pair=pair):
            got = run(m_new)
        errs = {"logits": rel_err(got[0], want[0]), "infonce": rel_err(got[1], want[1]), "ct": rel_err(got[2], want[2])}
        gerr = {k: norm_err(dict(m_new.named_parameters())[k].grad, g_ref[k]) for k in names}
        report("mm_model_dropin", task, act, {k: "%.1e" % v for k, v in errs.items()}, {k: "%.1e" % v for k, v in gerr.items()})
        assert max(errs.values()) < ltol, errs
        assert max(gerr.values()) < gtol, gerr
