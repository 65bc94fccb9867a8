"""CPU: the oracle restatement reproduces every golden fixture generated from the
reference's own files (oracle/make_golden.py)."""
import torch

from conftest import load_golden, rel_err
from oracle import restate
from oracle.detw import det_state_dict


def _params(g):
    return {k[2:]: v.clone().requires_grad_(True) for k, v in g.items() if k.startswith("w.")}


def test_pair_bias_golden():
    for tag in ("init", "pre"):
        g = load_golden("pair_bias_" + tag)
        p = _params(g)
        out = restate.pair_bias(g["in.dist"], g["in.edge_type"], p).view(g["out.bias"].shape)
        assert rel_err(out, g["out.bias"]) < 5e-6
        (out * g["in.upstream"]).sum().backward()
        for k, v in g.items():
            if k.startswith("grad."):
                assert rel_err(p[k[5:]].grad, v) < 5e-5, k


def test_encoder_golden():
    for tag in ("small", "nopad"):
        g = load_golden("encoder_" + tag)
        H, D, Fd, nl = [int(v) for v in g["cfg"]]
        p = _params(g)
        pm = g["in.tokens"].eq(0)
        pm = pm if pm.any() else None
        emb = g["in.emb"].clone().requires_grad_(True)
        b0 = g["in.bias"].clone().requires_grad_(True)
        bw = b0 * 1.0
        x, pair, delta, xn, dn = restate.encoder_with_pair(emb, bw, pm, p, H, nl)
        assert rel_err(x, g["out.x"]) < 5e-6
        assert rel_err(pair, g["out.pair"]) < 5e-6
        assert rel_err(delta, g["out.delta"]) < 5e-6
        assert torch.equal(torch.isinf(bw), torch.isinf(g["out.bias_after"]))      # Q1, bit-exact
        (x * g["in.up_x"]).sum().add((delta * g["in.up_delta"]).sum()).add(xn).add(dn).backward()
        assert rel_err(emb.grad, g["grad.emb"]) < 5e-5
        assert rel_err(b0.grad, g["grad.bias"]) < 5e-5


def test_encoder_slice_golden():
    g = load_golden("encoder_slice")
    H, D, Fd, nl, seed = [int(v) for v in g["cfg"]]
    from tests_util import slice_shapes
    p = {k: v.requires_grad_(True) for k, v in det_state_dict(slice_shapes(H, D, Fd, nl), seed=seed).items()}
    rep = restate.unimol_encoder(g["in.tokens"], g["in.dist"], g["in.edge_type"], p, heads=H, n_layers=nl)
    assert rel_err(rep, g["out.rep"]) < 5e-6
    (rep * g["in.up"]).sum().backward()
    for k, v in g.items():
        if k.startswith("grad."):
            assert rel_err(p[k[5:]].grad, v) < 5e-5, k


def test_infonce_golden():
    for tag in ("n16", "n37", "n64d512"):
        g = load_golden("infonce_" + tag)
        q, k = g["in.q"].clone().requires_grad_(True), g["in.k"].clone().requires_grad_(True)
        loss = restate.info_nce(q, k, 0.1)
        loss.backward()
        assert rel_err(loss, g["out.loss"]) < 1e-6
        assert rel_err(q.grad, g["grad.q"]) < 1e-5 and rel_err(k.grad, g["grad.k"]) < 1e-5


def test_ct_golden():
    for n in (16, 45):
        g = load_golden("ct_n%d" % n)
        f = g["in.feature"]
        cases = {
            "regress_w": lambda x: restate.ct_regress(x, g["in.y"], g["in.yhat"], weights=g["in.weights"], w=0.2),
            "regress_now": lambda x: restate.ct_regress(x, g["in.y"], g["in.yhat"], w=0.2),
            "single": lambda x: restate.ct_single(x, g["in.cls"]),
            "multi": lambda x: restate.ct_multi(x, g["in.multi"]),
        }
        for name, fn in cases.items():
            x = f.clone().requires_grad_(True)
            loss = fn(x)
            loss.backward()
            assert rel_err(loss, g["out." + name]) < 1e-6, name
            assert rel_err(x.grad, g["grad." + name]) < 1e-5, name
        for mode, lab in (("regress", "in.y"), ("single", "in.cls"), ("multi", "in.multi")):
            pos, neg, _ = restate.ct_masks(mode, g[lab], g["in.yhat"], 0.2)
            assert torch.equal(pos, g["out.%s_pos" % mode].bool())          # bit-exact
            assert torch.equal(neg, g["out.%s_neg" % mode].bool())


def test_fds_golden():
    g = load_golden("fds")
    cfg = dict(min_value=float(g["cfg.min_value"]), bin_width=float(g["cfg.bin_width"]), bucket_num=12,
               bucket_start=0, start_update=0, start_smooth=1, momentum=0.9)
    win = restate.fds_kernel_window("gaussian", 5, 1)
    assert rel_err(win, g["cfg.window"]) < 1e-6
    assert torch.equal(restate.fds_label_bins(g["in.labels"], cfg["min_value"], cfg["bin_width"]), g["out.bins"].long())
    nb, D = 12, 16
    st = {"epoch": torch.zeros(1), "running_mean": torch.zeros(nb, D), "running_var": torch.ones(nb, D),
          "running_mean_last_epoch": torch.zeros(nb, D), "running_var_last_epoch": torch.ones(nb, D),
          "smoothed_mean_last_epoch": torch.zeros(nb, D), "smoothed_var_last_epoch": torch.ones(nb, D),
          "num_samples_tracked": torch.zeros(nb)}
    restate.fds_update_running_stats(g["in.feats_e0"], g["in.labels"], 0, st, cfg)
    restate.fds_update_last_epoch_stats(1, st, win)
    for k in ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"):
        assert rel_err(st[k], g["out.e1." + k]) < 1e-5, k
    x = g["in.smooth_x"].clone().requires_grad_(True)
    xs = restate.fds_smooth(x * 1.0, g["in.labels"], 1, st, cfg)
    (xs * g["in.smooth_up"]).sum().backward()
    assert rel_err(xs, g["out.smooth"]) < 1e-5
    assert rel_err(x.grad, g["grad.smooth_x"]) < 1e-5
    sub = g["in.sub"].bool()
    xs2 = restate.fds_smooth(g["in.smooth_x"][sub].clone(), g["in.labels"][sub], 1, st, cfg)
    assert rel_err(xs2, g["out.smooth_noedge"]) < 1e-5
    restate.fds_update_running_stats(g["in.feats_e1"], g["in.labels"], 1, st, cfg)
    restate.fds_update_last_epoch_stats(2, st, win)
    for k in ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked"):
        assert rel_err(st[k], g["out.e2." + k]) < 1e-5, k
    m = g["in.cal_mat"]
    v1z = g["in.cal_v1"].clone()
    v1z[[2, 7]] = 0
    args = (g["in.cal_m1"], g["in.cal_v1"], g["in.cal_m2"], g["in.cal_v2"])
    assert rel_err(restate.calibrate_mean_var(m.clone(), *args), g["out.cal_full"]) < 1e-6
    assert rel_err(restate.calibrate_mean_var(m.clone(), args[0], v1z, args[2], args[3]), g["out.cal_zero_cols"]) < 1e-6
    assert rel_err(restate.calibrate_mean_var(m.clone(), args[0], args[1] * 0, args[2], args[3]), g["out.cal_tiny"]) < 1e-6


def test_featurise_restatement_matches_reference_fixture():
    """data/conformer.py coords2unimol + utils/util.py padding, executed from the reference tree by make_golden.py"""
    g = load_golden("featurise")
    dist, et = restate.featurise(g["in.src_tokens"], g["in.src_coord"], n_dict=31, pad_idx=0)
    assert torch.equal(dist, g["out.src_distance"])
    assert torch.equal(et, g["out.src_edge_type"])


def test_cross_modal_golden():
    """oracle/restate.py:cross_modal + fuse_pool against the fixture made by the reference's CrossAttentionModel."""
    g = load_golden("cross_modal")
    H, D, Fd, seed, ROWS = [int(v) for v in g["cfg"]]
    from tests_util import cross_layer_shapes
    shapes = {}
    for side in ("text_attention", "graph_attention"):
        shapes.update(cross_layer_shapes(D, Fd, side + ".layer.0."))
    p = {k: v.requires_grad_(True) for k, v in det_state_dict(shapes, seed=seed, std=0.05).items()}
    x1, x2 = g["in.x1"].clone().requires_grad_(True), g["in.x2"].clone().requires_grad_(True)
    t2g, g2t = restate.cross_modal(x1, x2, g["in.m1"], g["in.m2"], p, heads=H)
    assert rel_err(t2g, g["out.t2g"]) < 5e-6 and rel_err(g2t, g["out.g2t"]) < 5e-6
    pooled = restate.fuse_pool(t2g, g2t, g["in.m1"], g["in.m2"])
    assert rel_err(pooled, g["out.pooled"]) < 5e-6
    (pooled * g["in.up"]).sum().backward()
    assert rel_err(x1.grad, g["grad.x1"]) < 5e-5 and rel_err(x2.grad, g["grad.x2"]) < 5e-5
    for k, v in g.items():
        if k.startswith("grad.") and k not in ("grad.x1", "grad.x2") and not k.endswith("key.bias"):
            got = p[k[5:]].grad
            assert rel_err(got[:ROWS] if got.dim() == 2 else got, v) < 5e-5, k


def test_chemberta_golden():
    """oracle/restate.py:roberta_encoder against the fixture made by Hugging Face's RobertaModel (the reference's second-modality
    encoder, models/mm_model.py:475,562)."""
    g = load_golden("chemberta")
    H, D, Fd, nl, V, P, seed, ROWS = [int(v) for v in g["cfg"]]
    from tests_util import cross_layer_shapes
    shapes = {"embeddings.word_embeddings.weight": (V, D), "embeddings.token_type_embeddings.weight": (2, D),
              "embeddings.position_embeddings.weight": (P, D), "embeddings.LayerNorm.weight": (D,), "embeddings.LayerNorm.bias": (D,),
              "pooler.dense.weight": (D, D), "pooler.dense.bias": (D,)}
    for i in range(nl):
        shapes.update(cross_layer_shapes(D, Fd, "encoder.layer.%d." % i))
    p = {k: v.requires_grad_(True) for k, v in det_state_dict(shapes, seed=seed, std=0.05).items()}
    out = restate.roberta_encoder(g["in.ids"], g["in.mask"], p, heads=H, n_layers=nl)
    m = g["in.mask"].bool()
    assert rel_err(out[m], g["out.hidden"][m]) < 5e-6
    (out * g["in.up"]).sum().backward()
    for k, v in g.items():
        if k.startswith("grad."):
            got = p[k[5:]].grad
            got = got[:ROWS] if (got.dim() == 2 and "embeddings" not in k) else got
            assert rel_err(got, v) < 5e-5, k
