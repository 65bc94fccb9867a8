"""CPU: the host-side mirror of the reference interface -- signatures, defaults, state_dict names,
error behaviour, the task losses and the non-fused info_nce branches against fixtures produced by the reference's own
files (oracle/make_golden.py:gold_loss), the synthetic-molecule batch format, and the host arithmetic of the C ABI
(leading dimension of the pair tensor).  No kernel is launched here."""
import inspect
import os

import numpy as np
import pytest
import torch

import mmdti_b200  # noqa: F401
from mmdti_b200 import _lib, data
from mmdti_b200.models import contrastive, infonce, loss
from mmdti_b200.models.encoder import GaussianLayer, NonLinearHead, UnimolEncoder, gaussian
from mmdti_b200.models.fds import FDS, kernel_window
from mmdti_b200.models.transformers import TransformerEncoderWithPair

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _defaults(fn):
    return {k: v.default for k, v in inspect.signature(fn).parameters.items() if v.default is not inspect.Parameter.empty}


def test_constructor_and_function_signatures_match_the_reference():
    # models/transformers.py:32-46
    assert _defaults(TransformerEncoderWithPair.__init__) == dict(
        encoder_layers=6, embed_dim=768, ffn_embed_dim=3072, attention_heads=8, emb_dropout=0.1, dropout=0.1,
        attention_dropout=0.1, activation_dropout=0.0, max_seq_len=256, activation_fn="gelu", post_ln=False,
        no_final_head_layer_norm=False)
    assert list(inspect.signature(TransformerEncoderWithPair.forward).parameters)[:4] == ["self", "emb", "attn_mask", "padding_mask"]
    # models/mm_model.py:211,226,97 (dup models/encoder.py)
    assert list(inspect.signature(gaussian).parameters) == ["x", "mean", "std"]
    assert _defaults(GaussianLayer.__init__) == dict(K=128, edge_types=1024)
    assert list(inspect.signature(NonLinearHead.__init__).parameters) == ["self", "input_dim", "out_dim", "activation_fn", "hidden"]
    assert list(inspect.signature(UnimolEncoder.forward).parameters)[:4] == ["self", "src_tokens", "src_distance", "src_edge_type"]
    # models/infonce.py:11,42
    assert _defaults(infonce.InfoNCE.__init__) == dict(temperature=0.1, reduction="mean", negative_mode="unpaired")
    d = _defaults(infonce.info_nce)
    assert (d["negative_keys"], d["temperature"], d["reduction"], d["negative_mode"]) == (None, 0.1, "mean", "unpaired")
    # models/contrastive.py:3,62,114
    d = _defaults(contrastive.CT_Regress)
    assert (d["weights"], d["w"], d["t"], d["e"]) == (None, 0.2, 0.07, 0.01)
    d = _defaults(contrastive.CT_Single)
    assert (d["w"], d["t"], d["e"], d["lamda"]) == (0.2, 0.07, 0.2, 1) and torch.equal(d["weights"], torch.tensor([1]))
    d = _defaults(contrastive.CT_Multi)
    assert (d["weights"], d["w"], d["t"], d["e"], d["coef"]) == (None, 0.2, 0.07, 0.2, 1)
    # models/fds.py:33-35
    assert _defaults(FDS.__init__) == dict(bucket_num=100, bucket_start=0, start_update=0, start_smooth=1, kernel="gaussian",
                                           ks=5, sigma=2, momentum=0.9, device="cuda")
    for name in ("smooth", "update_last_epoch_stats", "update_running_stats", "reset"):
        assert callable(getattr(FDS, name))


def test_state_dict_names_and_shapes_match_the_unimol_checkpoint_layout():
    """the Uni-Mol checkpoint is loaded strict=False (mm_model.py:514), so a renamed key would be dropped
    silently -- every name and shape is pinned here."""
    sd = UnimolEncoder().state_dict()
    want = {"embed_tokens.weight": (31, 512), "gbf.means.weight": (1, 128), "gbf.stds.weight": (1, 128),
            "gbf.mul.weight": (961, 1), "gbf.bias.weight": (961, 1),
            "gbf_proj.linear1.weight": (128, 128), "gbf_proj.linear1.bias": (128,),
            "gbf_proj.linear2.weight": (64, 128), "gbf_proj.linear2.bias": (64,),
            "encoder.emb_layer_norm.weight": (512,), "encoder.emb_layer_norm.bias": (512,),
            "encoder.final_layer_norm.weight": (512,), "encoder.final_layer_norm.bias": (512,)}
    for i in range(15):
        p = "encoder.layers.%d." % i
        want.update({p + "self_attn.in_proj.weight": (1536, 512), p + "self_attn.in_proj.bias": (1536,),
                     p + "self_attn.out_proj.weight": (512, 512), p + "self_attn.out_proj.bias": (512,),
                     p + "self_attn_layer_norm.weight": (512,), p + "self_attn_layer_norm.bias": (512,),
                     p + "fc1.weight": (2048, 512), p + "fc1.bias": (2048,), p + "fc2.weight": (512, 2048), p + "fc2.bias": (512,),
                     p + "final_layer_norm.weight": (512,), p + "final_layer_norm.bias": (512,)})
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    assert sd["embed_tokens.weight"][0].abs().max() == 0            # padding row (init_bert_params)
    sd = infonce.InfoNCE(512, 512).state_dict()
    assert sorted(sd) == sorted("info_proj_%s.%d.%s" % (m, i, w) for m in ("query", "positive") for i in (0, 2) for w in ("weight", "bias"))
    assert tuple(sd["info_proj_query.2.weight"].shape) == (50, 512)
    fds = FDS(feature_dim=8, raw_data=np.array([0.0, 1.0, 2.0]), col_data=None, using_scale=False, bucket_num=10, device="cpu")
    assert list(fds.state_dict()) == ["epoch", "running_mean", "running_var", "running_mean_last_epoch", "running_var_last_epoch",
                                      "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"]
    assert tuple(fds.running_mean.shape) == (10, 8)


def test_fds_kernel_window_matches_the_reference_taps():
    """models/fds.py:69-84: gaussian window = gaussian_filter1d of a unit impulse, normalised by ITS SUM (the reference
    divides by the sum, not the max, for the gaussian kernel); triang / laplace variants."""
    from scipy.ndimage import gaussian_filter1d
    from scipy.signal.windows import triang
    base = np.zeros(5)
    base[2] = 1.0
    w = gaussian_filter1d(base, sigma=2)
    np.testing.assert_allclose(kernel_window("gaussian", 5, 2), w / w.sum(), rtol=1e-6)
    np.testing.assert_allclose(kernel_window("triang", 5, 2), triang(5) / triang(5).sum(), rtol=1e-6)
    lap = np.array([np.exp(-abs(x) / 2.0) / 4.0 for x in range(-2, 3)])
    np.testing.assert_allclose(kernel_window("laplace", 5, 2), lap / lap.sum(), rtol=1e-6)


def test_info_nce_value_errors():
    """models/infonce.py:45-67: each shape error raises ValueError before any device work."""
    r = torch.randn
    bad = [((r(4), r(4, 3)), {}), ((r(4, 3), r(4)), {}), ((r(4, 3), r(5, 3)), {}), ((r(4, 3), r(4, 2)), {}),
           ((r(4, 3), r(4, 3), r(2, 2, 3)), dict(negative_mode="unpaired")),
           ((r(4, 3), r(4, 3), r(2, 3)), dict(negative_mode="paired")),
           ((r(4, 3), r(4, 3), r(3, 2, 3)), dict(negative_mode="paired")),
           ((r(4, 3), r(4, 3), r(2, 5)), dict(negative_mode="unpaired"))]
    for args, kw in bad:
        with pytest.raises(ValueError):
            infonce.info_nce(*args, **kw)


def test_task_losses_and_explicit_negative_info_nce_match_the_reference_fixture():
    z = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, "loss.npz")).items()}
    x, y, yb, yb_nan, y_nan = z["in.x"], z["in.y"], z["in.yb"], z["in.yb_nan"], z["in.y_nan"]

    def close(got, key, tol=1e-6):
        torch.testing.assert_close(got.float(), z[key].float(), rtol=tol, atol=tol, msg=lambda m: key + ": " + m)

    close(loss.RMSELoss()(x, y), "out.rmse")
    ghmc, ghmr = loss.GHMC_Loss(bins=10, alpha=0.5), loss.GHMR_Loss(bins=10, alpha=0.5, mu=0.02)
    close(ghmc(x, yb), "out.ghmc_1")
    close(ghmc(0.5 * x, yb), "out.ghmc_2")          # second call: EMA of the bin populations (loss.py:84-86)
    close(ghmr(x, y), "out.ghmr_1")
    close(ghmr(0.5 * x, y), "out.ghmr_2")
    close(loss.MaskedBCEWithLogitsLoss()(x, yb_nan), "out.masked_bce")
    close(loss.MAEwithNan(x, y_nan), "out.mae_nan")
    close(loss.BCEwithNan(x, yb_nan), "out.bce_nan")
    close(loss.FocalLoss(z["in.prob"], yb[:, 0]), "out.focal")
    close(loss.FocalLossWithLogits(x, yb_nan), "out.focal_logits")
    close(loss.myCrossEntropyLoss(x, z["in.cls"]), "out.ce")
    # models/infonce.py:71-88 -- stock composition, outside the fused path; runs wherever the tensors live
    q, k = z["in.q"], z["in.k"]
    close(infonce.info_nce(q, k, z["in.neg_unpaired"], temperature=0.1, negative_mode="unpaired"), "out.nce_unpaired")
    close(infonce.info_nce(q, k, z["in.neg_paired"], temperature=0.2, negative_mode="paired"), "out.nce_paired")
    close(infonce.info_nce(q, k, temperature=0.1, reduction="none"), "out.nce_none")
    # the reference's symmetric form only runs with M + 1 == N explicit negatives; anything else raises (infonce.py:98)
    with pytest.raises(ValueError):
        infonce.info_nce(q, k, z["in.neg_unpaired"][:5])


def test_fused_ops_refuse_cpu_tensors():
    """No CPU fallback: the fused paths raise on host tensors instead of computing something else."""
    from mmdti_b200._lib import MMDTIError
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    f, yv = torch.randn(8, 16), torch.randn(8, 1)
    for fn in (lambda: infonce.info_nce(f, f.clone()), lambda: contrastive.CT_Regress(f, yv, yv.clone()),
               lambda: contrastive.CT_Single(f, torch.randint(0, 2, (8, 1)), None),
               lambda: data.featurise(torch.ones(2, 5, dtype=torch.long), torch.zeros(2, 5, 3))):
        with pytest.raises((MMDTIError, RuntimeError, AssertionError)):
            fn()


def test_synthetic_molecules_follow_the_reference_batch_format():
    """[CLS]=1, atoms in 4..29, [SEP]=2, pad 0; edge type = tok_i * 31 + tok_j with padded pairs 0
    (data/conformer.py:204-218, utils/util.py:41-105); distances symmetric with a zero diagonal."""
    tokens, dist, et, coord = data.synthetic_molecules(6, 20, seed=3, ragged=True)
    B, L = tokens.shape
    assert L == 22 and dist.shape == (B, L, L) and et.shape == (B, L, L) and et.dtype == torch.int64
    lens = (tokens != 0).sum(1)
    assert lens.min() < L                                            # ragged
    for b in range(B):
        n = int(lens[b])
        assert tokens[b, 0] == 1 and tokens[b, n - 1] == 2 and (tokens[b, n:] == 0).all()
        assert ((tokens[b, 1:n - 1] >= 4) & (tokens[b, 1:n - 1] <= 29)).all()
        want = tokens[b, :n, None] * 31 + tokens[b, None, :n]
        assert torch.equal(et[b, :n, :n], want) and (et[b, n:] == 0).all() and (et[b, :, n:] == 0).all()
        assert (dist[b, n:] == 0).all() and (dist[b, :, n:] == 0).all()
    assert torch.equal(dist, dist.transpose(1, 2)) and (dist.diagonal(dim1=1, dim2=2) == 0).all()
    # same seed, same batch
    again = data.synthetic_molecules(6, 20, seed=3, ragged=True)
    assert all(torch.equal(a, b) for a, b in zip((tokens, dist, et), again[:3]))


def test_padding_helpers_match_the_reference_rules():
    """utils/util.py:7-105: right padding to the batch maximum rounded up to a multiple of 8 when asked."""
    toks = [torch.tensor([1, 5, 2]), torch.tensor([1, 5, 6, 7, 2])]
    out = data.pad_1d_tokens(toks, 0)
    assert out.tolist() == [[1, 5, 2, 0, 0], [1, 5, 6, 7, 2]]
    assert data.pad_1d_tokens(toks, 0, pad_to_multiple=8).shape == (2, 8)
    assert data.pad_1d_tokens(toks, 0, left_pad=True)[0].tolist() == [0, 0, 1, 5, 2]
    m = data.pad_2d([torch.ones(3, 3), torch.ones(5, 5)], 0)
    assert m.shape == (2, 5, 5) and m[0].sum() == 9 and m[0, :3, :3].sum() == 9
    c = data.pad_coords([torch.ones(3, 3), torch.ones(5, 3)], 0.0)
    assert c.shape == (2, 5, 3) and c[0, 3:].abs().sum() == 0


def test_pair_tensor_leading_dimension_rule():
    """Lp = 8 * NKB >= L with NKB odd (16-byte rows; Lp = 8 mod 16 keeps the slab bank-conflict-free),
    NKB from the instantiated set {3, 5, 9, 13, 17, 25, 33}; L > 264 is refused (-1), never truncated.
    Host arithmetic of the C ABI -- no device needed."""
    lib = _lib.lib()
    buckets = [8 * n for n in (3, 5, 9, 13, 17, 25, 33)]
    for L in range(1, 265):
        lp = lib.mmdti_pair_ld(L)
        assert lp == min(b for b in buckets if b >= L) and lp % 16 == 8, (L, lp)
    assert lib.mmdti_pair_ld(66) == 72 and lib.mmdti_pair_ld(258) == 264
    for L in (265, 512):
        assert lib.mmdti_pair_ld(L) == -1


def test_chemberta_encoder_dispatch_and_checkpoint_loading(tmp_path):
    """models/encoder.py:548-572 mirror: configurations the fused layer covers get mmdti_b200's RobertaModel (same state_dict names
    as Hugging Face's, checkpoints load strictly), anything else the stock HF module with a warning."""
    import warnings
    import torch
    from transformers import RobertaConfig, RobertaModel
    from mmdti_b200.models import chemberta
    from mmdti_b200.models.encoder import ChembertaEncoder
    cfg = RobertaConfig(vocab_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=4, intermediate_size=256,
                        max_position_embeddings=40, pad_token_id=1)
    hf = RobertaModel(cfg)
    hf.save_pretrained(str(tmp_path))
    enc = ChembertaEncoder(str(tmp_path))
    assert isinstance(enc.bert, chemberta.RobertaModel)
    want = hf.state_dict()
    got = enc.bert.state_dict()
    assert sorted(got) == sorted(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    big = RobertaConfig(vocab_size=64, hidden_size=768, num_hidden_layers=1, num_attention_heads=12, intermediate_size=3072,
                        max_position_embeddings=40, pad_token_id=1)
    assert not chemberta.supported(big)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        stock = ChembertaEncoder(big)
    assert type(stock.bert).__module__.startswith("transformers") and any("stock HF" in str(x.message) for x in w)


def test_cross_modal_dropin_has_the_reference_state_dict_and_config():
    """models/mm_model.py:361-406: the drop-in fusion block exposes exactly the parameter names and shapes of the reference's
    CrossAttentionModel, and crossmodal_config() the reference's values (needs a copy of the reference tree)."""
    import pytest
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("no copy of the reference tree on this machine")
    mm = ref_loader.load()["mm_model"]
    from mmdti_b200.models.cross_modal import CrossAttentionModel, crossmodal_config
    ref_cfg, cfg = mm.crossmodal_config(), crossmodal_config()
    for k, v in vars(cfg).items():
        assert getattr(ref_cfg, k) == v, k
    want = {k: tuple(v.shape) for k, v in mm.CrossAttentionModel(ref_cfg, num_layers=2).state_dict().items()}
    got = {k: tuple(v.shape) for k, v in CrossAttentionModel(cfg, num_layers=2).state_dict().items()}
    assert got == want
