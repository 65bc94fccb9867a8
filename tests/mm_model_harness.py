"""Builds the reference's OWN ``MM_Model`` (models/mm_model.py:408-618) twice -- once with the reference's classes (Uni-Core
from oracle/shims), once with the mmdti_b200 drop-in modules patched into the reference module namespace -- with the same
random-init weights and a small synthetic ChemBERTa, so that a test can compare ``logits, ct_loss, rnc_loss`` and gradients
(SURVEY.md §4 test iv).  TEST INFRASTRUCTURE: needs a copy of the reference tree (``oracle.ref_loader.available()``)."""
import os
import tempfile

import torch

MOL_DICT = ["[PAD]", "[CLS]", "[SEP]", "[UNK]", "C", "N", "O", "S", "H", "Cl", "F", "Br", "I", "Si", "P", "B", "Na", "K", "Al", "Ca", "Sn",
            "As", "Hg", "Fe", "Zn", "Cr", "Se", "Gd", "Au", "Li"]


def make_assets(tmp):
    """mol.dict.txt + an (empty) Uni-Mol checkpoint + a random-init 2-layer RoBERTa standing in for ChemBERTa (hidden 512)."""
    from transformers import RobertaConfig, RobertaModel
    with open(os.path.join(tmp, "mol.dict.txt"), "w") as fh:
        fh.write("\n".join(MOL_DICT) + "\n")
    torch.save({"model": {}}, os.path.join(tmp, "unimol.pt"))
    torch.manual_seed(3)
    cfg = RobertaConfig(vocab_size=64, hidden_size=512, num_hidden_layers=2, num_attention_heads=8, intermediate_size=1024,
                        max_position_embeddings=80, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pad_token_id=1)
    bert_dir = os.path.join(tmp, "chemberta")
    RobertaModel(cfg).save_pretrained(bert_dir)
    return os.path.join(tmp, "unimol.pt"), bert_dir


def build(ref, dropin, task="regression", layers=2):
    """ref: dict of reference modules (oracle.ref_loader.load()).  dropin=True patches the mmdti_b200 modules into the reference's
    module namespaces for the duration of the constructor.  Returns the model (all dropout probabilities set to 0)."""
    import transformers
    mm = ref["mm_model"]
    tmp = tempfile.mkdtemp(prefix="mmdti_mm_")
    unimol_pt, bert_dir = make_assets(tmp)
    saved = {}

    def patch(mod, name, val):
        saved[(mod, name)] = getattr(mod, name)
        setattr(mod, name, val)

    patch(transformers.AutoTokenizer, "from_pretrained", staticmethod(lambda *a, **k: None))      # the tokenizer is not used by forward
    arch = mm.molecule_architecture
    patch(mm, "molecule_architecture", lambda: _with(arch(), encoder_layers=layers))
    if dropin:
        from mmdti_b200.models import contrastive as dct
        from mmdti_b200.models import cross_modal as dcm
        from mmdti_b200.models import encoder as denc
        from mmdti_b200.models import fds as dfds
        from mmdti_b200.models import infonce as dinf
        from mmdti_b200.models import transformers as dtr
        patch(mm, "BACKBONE", {"transformer": dtr.TransformerEncoderWithPair})
        patch(mm, "GaussianLayer", denc.GaussianLayer)
        patch(mm, "NonLinearHead", denc.NonLinearHead)
        patch(mm, "InfoNCE", dinf.InfoNCE)
        patch(mm, "FDS", dfds.FDS)
        patch(mm, "CrossAttentionModel", dcm.CrossAttentionModel)
        from mmdti_b200.models import chemberta as dcb
        patch(mm, "AutoModel", dcb.RobertaModel)             # models/mm_model.py:475: AutoModel.from_pretrained(chemberta_dir)
        for n in ("CT_Regress", "CT_Single", "CT_Multi"):
            patch(ref["contrastive"], n, getattr(dct, n))
    try:
        torch.manual_seed(0)
        model = mm.MM_Model(output_dim=1, task=task, fds=False, chemberta_dir=bert_dir, unimol_dir=unimol_pt, ct_w=0.2)
    finally:
        for (mod, name), val in saved.items():
            setattr(mod, name, val)
    zero_dropout(model)
    return model


def _with(args, **kw):
    for k, v in kw.items():
        setattr(args, k, v)
    return args


def zero_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        for attr in ("dropout", "attention_dropout", "activation_dropout", "emb_dropout", "embed_dropout", "pooler_dropout"):
            if isinstance(getattr(m, attr, None), float):
                setattr(m, attr, 0.0)


def batch(B=4, n_atoms=12, S=10, seed=1):
    from mmdti_b200.data import synthetic_molecules
    tokens, dist, et, _ = synthetic_molecules(B, n_atoms, seed=seed, ragged=True, n_dict=len(MOL_DICT) + 1)
    g = torch.Generator().manual_seed(seed + 5)
    ids = torch.randint(4, 60, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    am[1, S - 3:] = 0
    ids[1, S - 3:] = 1
    y = torch.randn(B, 1, generator=g)
    w = torch.rand(B, generator=g) + 0.5
    return dict(src_tokens=tokens, src_distance=dist, src_edge_type=et, input_ids=ids, attention_mask=am), y, w / w.mean()
