"""GPU: the whole hot path chained the way MM_Model.forward chains it (models/mm_model.py:545-591):
conformer encoder -> InfoNCE against a second modality -> masked mean pooling -> FDS.smooth (in place;
the CT loss sees the smoothed features, Q11) -> regression head -> ConR, loss = task + 0.1 infonce + 0.1 ct
(tasks/trainer.py:68-69,193), forward + backward, against the same chain built from the oracle restatement.
ChemBERTa and the cross-modal block are outside the path: the second modality is a fixed random tensor."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mmdti_b200
from conftest import norm_err, rel_err
from mmdti_b200.data import synthetic_molecules
from mmdti_b200.models.contrastive import CT_Regress
from mmdti_b200.models.encoder import UnimolEncoder
from mmdti_b200.models.fds import FDS
from mmdti_b200.models.infonce import InfoNCE
from oracle import restate

from tolerances import TOL

pytestmark = pytest.mark.gpu


def _fds_state(nb, D, gen):
    return {"running_mean": torch.randn(nb, D, generator=gen) * 0.1, "running_var": torch.rand(nb, D, generator=gen) + 0.5,
            "smoothed_mean": torch.randn(nb, D, generator=gen) * 0.1, "smoothed_var": torch.rand(nb, D, generator=gen) + 0.5}


@pytest.mark.parametrize("mode,ltol,gtol", [("fp32",) + TOL["step.fp32"], ("bf16",) + TOL["step.bf16"]])
def test_hot_path_training_step_matches_oracle(mode, ltol, gtol, report):
    B, n_atoms, S, nl, nb = 12, 20, 9, 2, 10
    gen = torch.Generator().manual_seed(77)
    torch.manual_seed(5)
    enc = UnimolEncoder(encoder_layers=nl).cuda().eval()          # eval: no dropout, FDS / CT still active below
    inf = InfoNCE(512, 512).cuda().eval()
    head = torch.nn.Linear(512, 1).cuda()
    tokens, dist, et, _ = synthetic_molecules(B, n_atoms, seed=3, ragged=True)
    smiles = torch.randn(B, S, 512, generator=gen) * 0.5
    y = torch.randn(B, 1, generator=gen)
    wts = torch.rand(B, generator=gen) + 0.5
    st = _fds_state(nb, 512, gen)
    cfg = dict(min_value=-2.0, bin_width=0.4, bucket_num=nb, bucket_start=0, start_smooth=1)
    fds = FDS(feature_dim=512, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=nb).cuda()
    fds.min_value, fds.bin_width = cfg["min_value"], cfg["bin_width"]
    fds.running_mean_last_epoch.copy_(st["running_mean"]); fds.running_var_last_epoch.copy_(st["running_var"])
    fds.smoothed_mean_last_epoch.copy_(st["smoothed_mean"]); fds.smoothed_var_last_epoch.copy_(st["smoothed_var"])

    # ---- oracle chain (CPU, fp32)
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc.state_dict().items()}
    pi = {"infonce." + k: v.detach().cpu().clone().requires_grad_(True) for k, v in inf.state_dict().items()}
    hw, hb = head.weight.detach().cpu().clone().requires_grad_(True), head.bias.detach().cpu().clone().requires_grad_(True)
    rep = restate.unimol_encoder(tokens, dist, et, p, heads=64, n_layers=nl)
    l_inf = restate.infonce_head(rep, smiles, pi)
    mask = tokens.ne(0).float().unsqueeze(-1)
    pooled = (rep * mask).sum(1) / mask.sum(1)
    ost = {"running_mean_last_epoch": st["running_mean"], "running_var_last_epoch": st["running_var"],
           "smoothed_mean_last_epoch": st["smoothed_mean"], "smoothed_var_last_epoch": st["smoothed_var"]}
    feats = restate.fds_smooth(pooled * 1.0, y, 1, ost, cfg)
    logits = F.linear(feats, hw, hb)
    l_ct = restate.ct_regress(feats, y, logits, weights=wts, w=0.2)
    loss_ref = F.mse_loss(logits, y) + 0.1 * l_inf + 0.1 * l_ct
    loss_ref.backward()

    # ---- drop-in chain (GPU)
    with mmdti_b200.precision(act=mode):
        rep_g = enc(tokens.cuda(), dist.cuda(), et.cuda()).float()
        li = inf(rep_g, smiles.cuda())
        mk = tokens.ne(0).float().unsqueeze(-1).cuda()
        pooled_g = (rep_g * mk).sum(1) / mk.sum(1)
        feats_g = fds.smooth(pooled_g * 1.0, y.cuda(), 1)
        logits_g = head(feats_g)
        lc = CT_Regress(feats_g, y.cuda(), logits_g, weights=wts.cuda(), w=0.2)
        loss = F.mse_loss(logits_g, y.cuda()) + 0.1 * li + 0.1 * lc
        loss.backward()
    torch.cuda.synchronize()
    errs = {"loss": rel_err(loss, loss_ref), "infonce": rel_err(li, l_inf), "ct": rel_err(lc, l_ct),
            "rep": norm_err(rep_g, rep)}
    named = dict(enc.named_parameters())
    for k in ("encoder.layers.0.self_attn.in_proj.weight", "encoder.layers.1.fc2.weight", "gbf.means.weight", "embed_tokens.weight"):
        errs["d_" + k] = norm_err(named[k].grad, p[k].grad)
    errs["d_head"] = norm_err(head.weight.grad, hw.grad)
    errs["d_infonce_q0"] = norm_err(inf.info_proj_query[0].weight.grad, pi["infonce.info_proj_query.0.weight"].grad)
    report("hot_path_step", mode, {k: "%.1e" % v for k, v in errs.items()})
    assert errs["loss"] < ltol and errs["infonce"] < ltol and errs["ct"] < ltol, errs
    assert all(v < gtol for k, v in errs.items() if k.startswith("d_") or k == "rep"), errs
