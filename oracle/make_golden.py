"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FILES (read-only, from
/root/reference, via oracle/ref_loader.py) on seeded inputs.  TEST INFRASTRUCTURE.

Run in the build container only:   python -m oracle.make_golden
It also asserts that oracle/restate.py reproduces every fixture (fp32, tight tolerance)
— that is what pins the restatement.  The fixtures hold inputs, weights (by state_dict
name), outputs and gradients, so the tests never depend on RNG reproducibility.

Pinned by the reference's own code: gaussian/GaussianLayer/NonLinearHead
(models/mm_model.py), TransformerEncoderWithPair.forward (models/transformers.py),
InfoNCE/info_nce (models/infonce.py), CT_Regress/CT_Single/CT_Multi
(models/contrastive.py), FDS (models/fds.py) + calibrate_mean_var (utils/util.py).
NOT pinned (third-party Uni-Core restated in oracle/shims/unicore): the arithmetic inside
TransformerEncoderLayer.
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, restate  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def _save(name, d):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(d))
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in _np(d).items() if not k.startswith("w.")})


def _close(a, b, tol, what):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), what + ": finiteness pattern differs"
    assert torch.equal(a[~fin], b[~fin]) or (~fin).sum() == 0, what + ": inf pattern differs"
    err = (a[fin] - b[fin]).abs().max().item() if fin.any() else 0.0
    scale = max(b[fin].abs().max().item(), 1e-30) if fin.any() else 1.0
    assert err <= tol * max(scale, 1.0), "%s: max err %.3e (scale %.3e)" % (what, err, scale)


def synth_molecules(B, n_atoms_max, seed, ragged=True, n_dict=31):
    """Synthetic conformers in the input format of data/conformer.py:204-212."""
    g = torch.Generator().manual_seed(seed)
    L = n_atoms_max + 2
    tokens = torch.zeros(B, L, dtype=torch.long)
    dist = torch.zeros(B, L, L)
    et = torch.zeros(B, L, L, dtype=torch.long)
    for b in range(B):
        n = n_atoms_max if (not ragged or b == 0) else int(torch.randint((n_atoms_max + 1) // 2, n_atoms_max + 1, (1,), generator=g))
        t = torch.cat([torch.tensor([1]), torch.randint(4, 30, (n,), generator=g), torch.tensor([2])])
        xyz = torch.randn(n, 3, generator=g, dtype=torch.float64) * 2.0
        xyz = xyz - xyz.mean(0, keepdim=True)
        xyz = torch.cat([torch.zeros(1, 3, dtype=torch.float64), xyz, torch.zeros(1, 3, dtype=torch.float64)])
        d = (xyz[:, None, :] - xyz[None, :, :]).pow(2).sum(-1).sqrt().float()
        m = n + 2
        tokens[b, :m] = t
        dist[b, :m, :m] = d
        et[b, :m, :m] = t[:, None] * n_dict + t[None, :]
    return tokens, dist, et


def gold_pair_bias(ref):
    mm = ref["mm_model"]
    for tag, pretrained_like in (("init", False), ("pre", True)):
        torch.manual_seed(11 if pretrained_like else 10)
        gbf = mm.GaussianLayer(128, 961)
        proj = mm.NonLinearHead(128, 64, "gelu")
        if not pretrained_like:                 # what init_bert_params leaves (Q6)
            for m_ in list(gbf.modules()) + list(proj.modules()):
                sys.modules["unicore.modules"].init_bert_params(m_)
        else:
            with torch.no_grad():
                gbf.mul.weight.add_(0.05 * torch.randn_like(gbf.mul.weight))
                gbf.bias.weight.add_(0.05 * torch.randn_like(gbf.bias.weight))
                proj.linear1.weight.normal_(0, 0.08)
                proj.linear2.weight.normal_(0, 0.08)
                proj.linear1.bias.normal_(0, 0.05)
                proj.linear2.bias.normal_(0, 0.05)
        tokens, dist, et = synth_molecules(3, 9, seed=100)
        o = proj(gbf(dist, et))
        bias = o.permute(0, 3, 1, 2).contiguous()
        up = torch.randn(bias.shape, generator=torch.Generator().manual_seed(5))
        (bias * up).sum().backward()
        params = {"gbf." + k: v for k, v in gbf.state_dict().items()}
        params.update({"gbf_proj." + k: v for k, v in proj.state_dict().items()})
        grads = {"gbf." + k: v.grad for k, v in gbf.named_parameters()}
        grads.update({"gbf_proj." + k: v.grad for k, v in proj.named_parameters()})
        d = {"in.tokens": tokens, "in.dist": dist, "in.edge_type": et, "in.upstream": up,
             "out.bias": bias}
        d.update({"w." + k: v for k, v in params.items()})
        d.update({"grad." + k: v for k, v in grads.items()})
        # pin the restatement
        p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        mine = restate.pair_bias(dist, et, p).view(bias.shape)
        _close(mine, bias, 2e-6, "pair_bias." + tag)
        (mine * up).sum().backward()
        for k in grads:
            _close(p[k].grad, grads[k], 2e-5, "pair_bias.%s.grad.%s" % (tag, k))
        _save("pair_bias_" + tag, d)


def gold_encoder(ref):
    tr = ref["transformers"]
    for tag, H, D, Fd, nl, B, n_atoms, ragged in (("small", 8, 64, 128, 3, 3, 10, True),
                                                   ("nopad", 4, 32, 64, 2, 2, 6, False)):
        torch.manual_seed(20)
        enc = tr.TransformerEncoderWithPair(encoder_layers=nl, embed_dim=D, ffn_embed_dim=Fd,
                                            attention_heads=H, emb_dropout=0.1, dropout=0.1,
                                            attention_dropout=0.1, activation_dropout=0.0,
                                            max_seq_len=512, activation_fn="gelu",
                                            no_final_head_layer_norm=True)
        for m_ in enc.modules():
            sys.modules["unicore.modules"].init_bert_params(m_)
        with torch.no_grad():                   # make LN affine and weights non-trivial
            for n_, p_ in enc.named_parameters():
                if "layer_norm" in n_:
                    p_.add_(0.1 * torch.randn_like(p_))
                elif n_.endswith("bias"):
                    p_.normal_(0, 0.05)
                else:
                    p_.normal_(0, 0.15)
        enc.eval()                              # dropout off: RNG streams cannot match
        tokens, _, _ = synth_molecules(B, n_atoms, seed=200, ragged=ragged)
        L = tokens.shape[1]
        pm = tokens.eq(0)
        pm = pm if pm.any() else None
        g = torch.Generator().manual_seed(21)
        emb = torch.randn(B, L, D, generator=g).requires_grad_(True)
        bias0 = torch.randn(B * H, L, L, generator=g)
        bias_in = bias0.clone().requires_grad_(True)
        bias_work = bias_in * 1.0               # non-leaf so the in-place fill is legal
        x, pair, delta, xn, dn = enc(emb, attn_mask=bias_work, padding_mask=pm)
        ux = torch.randn(x.shape, generator=g)
        up = torch.randn(delta.shape, generator=g)
        (x * ux).sum().add((delta * up).sum()).add(xn).add(dn).backward()
        d = {"in.tokens": tokens, "in.emb": emb, "in.bias": bias0, "in.up_x": ux, "in.up_delta": up,
             "cfg": np.array([H, D, Fd, nl]),
             "out.x": x, "out.pair": pair, "out.delta": delta, "out.x_norm": xn, "out.delta_norm": dn,
             "out.bias_after": bias_work,        # the mutated caller tensor (Q1)
             "grad.emb": emb.grad, "grad.bias": bias_in.grad}
        sd = {"encoder." + k: v for k, v in enc.state_dict().items()}
        d.update({"w." + k: v for k, v in sd.items()})
        gsel = ["layers.0.self_attn.in_proj.weight", "layers.0.self_attn.in_proj.bias",
                "layers.%d.fc2.weight" % (nl - 1), "layers.1.self_attn_layer_norm.weight",
                "emb_layer_norm.bias", "final_layer_norm.weight", "layers.0.self_attn.out_proj.weight"]
        named = dict(enc.named_parameters())
        for k in gsel:
            d["grad.encoder." + k] = named[k].grad
        # pin the restatement (of transformers.py; the layer itself is the shim's)
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        emb2 = emb.detach().clone().requires_grad_(True)
        b2 = bias0.clone().requires_grad_(True)
        bw = b2 * 1.0
        r = restate.encoder_with_pair(emb2, bw, pm, p, H, nl)
        for a, b_, nm in zip(r, (x, pair, delta, xn, dn), ("x", "pair", "delta", "xn", "dn")):
            _close(a, b_, 1e-6, "encoder.%s.%s" % (tag, nm))
        assert torch.equal(bw, bias_work), "in-place mask merge differs"
        (r[0] * ux).sum().add((r[2] * up).sum()).add(r[3]).add(r[4]).backward()
        _close(emb2.grad, emb.grad, 1e-5, "encoder.grad.emb")
        _close(b2.grad, bias_in.grad, 1e-5, "encoder.grad.bias")
        for k in gsel:
            _close(p["encoder." + k].grad, named[k].grad, 1e-5, "encoder.grad." + k)
        _save("encoder_" + tag, d)


def gold_encoder_slice(ref):
    """embed -> gbf -> gbf_proj -> encoder exactly as models/mm_model.py:545-559, with the
    reference classes at the production geometry (64 heads x 8, 512-d, FFN 2048), 2 layers.
    Weights come from oracle/detw.py (deterministic, not stored: 9 M parameters)."""
    from oracle.detw import det_state_dict
    mm, tr = ref["mm_model"], ref["transformers"]
    H, D, Fd, nl = 64, 512, 2048, 2
    emb_tok = torch.nn.Embedding(31, D, 0)
    gbf = mm.GaussianLayer(128, 961)
    proj = mm.NonLinearHead(128, H, "gelu")
    enc = tr.TransformerEncoderWithPair(encoder_layers=nl, embed_dim=D, ffn_embed_dim=Fd,
                                        attention_heads=H, max_seq_len=512,
                                        no_final_head_layer_norm=True)
    mods = torch.nn.ModuleDict({"embed_tokens": emb_tok, "gbf": gbf, "gbf_proj": proj, "encoder": enc})
    sd = det_state_dict({k: tuple(v.shape) for k, v in mods.state_dict().items()}, seed=3)
    mods.load_state_dict(sd)
    mods.eval()
    tokens, dist, et = synth_molecules(2, 12, seed=300)
    pm = tokens.eq(0)
    x = emb_tok(tokens)
    b = proj(gbf(dist, et)).permute(0, 3, 1, 2).contiguous()
    b = b.view(-1, b.size(-2), b.size(-1))
    rep = enc(x, padding_mask=pm, attn_mask=b)[0]
    g = torch.randn(rep.shape, generator=torch.Generator().manual_seed(31))
    (rep * g).sum().backward()
    named = dict(mods.named_parameters())
    gsel = ["embed_tokens.weight", "gbf.means.weight", "gbf.stds.weight", "gbf.mul.weight",
            "gbf.bias.weight", "gbf_proj.linear1.weight", "gbf_proj.linear2.bias",
            "encoder.layers.0.self_attn.in_proj.bias", "encoder.layers.1.fc1.bias",
            "encoder.layers.0.self_attn_layer_norm.weight"]
    d = {"in.tokens": tokens, "in.dist": dist, "in.edge_type": et, "in.up": g, "out.rep": rep,
         "cfg": np.array([H, D, Fd, nl, 3])}
    for k in gsel:
        d["grad." + k] = named[k].grad
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mine = restate.unimol_encoder(tokens, dist, et, p, heads=H, n_layers=nl)
    _close(mine, rep, 2e-6, "slice.rep")
    (mine * g).sum().backward()
    for k in gsel:
        _close(p[k].grad, named[k].grad, 2e-5, "slice.grad." + k)
    _save("encoder_slice", d)


def gold_encoder_15l(ref):
    """The configuration bench.py times (BASELINE configs[1] geometry): the reference's OWN classes — embed -> gbf ->
    gbf_proj -> TransformerEncoderWithPair exactly as models/mm_model.py:545-559 — at 15 layers, 64 heads x 8, 512-d,
    FFN 2048 (models/mm_model.py:325-343), L = 66 ragged, B = 4.  Weights: oracle/detw.py (47 M parameters, not
    stored).  Two weight regimes: 'init' (std 0.02 = init_bert_params) and 'wide' (std 0.05: sharper softmax, larger
    pair updates, i.e. what the 15-add pair chain looks like away from initialisation).
    Stored: all_repr, the residual stream of molecule 0 after EVERY layer (per-layer error growth), the pair tensor
    of molecule 0 / head 0 after every layer, 12 gradients (rows [:ROWS] of the big ones)."""
    from oracle.detw import det_state_dict
    mm, tr = ref["mm_model"], ref["transformers"]
    H, D, Fd, nl, ROWS = 64, 512, 2048, 15, 48
    for tag, std, seed in (("init", 0.02, 7), ("wide", 0.05, 8)):
        emb_tok = torch.nn.Embedding(31, D, 0)
        gbf = mm.GaussianLayer(128, 961)
        proj = mm.NonLinearHead(128, H, "gelu")
        enc = tr.TransformerEncoderWithPair(encoder_layers=nl, embed_dim=D, ffn_embed_dim=Fd,
                                            attention_heads=H, max_seq_len=512,
                                            no_final_head_layer_norm=True)
        mods = torch.nn.ModuleDict({"embed_tokens": emb_tok, "gbf": gbf, "gbf_proj": proj, "encoder": enc})
        sd = det_state_dict({k: tuple(v.shape) for k, v in mods.state_dict().items()}, seed=seed, std=std)
        mods.load_state_dict(sd)
        mods.eval()
        tokens, dist, et = synth_molecules(4, 64, seed=700 + seed)
        pm = tokens.eq(0)
        assert pm.any()
        per_layer_x, per_layer_pair = [], []

        def grab(_m, _i, o_):
            per_layer_x.append(o_[0][0].detach().clone())
            per_layer_pair.append(o_[1][0].detach().clone())

        hooks = [layer.register_forward_hook(grab) for layer in enc.layers]
        x = emb_tok(tokens)
        b = proj(gbf(dist, et)).permute(0, 3, 1, 2).contiguous()
        b = b.view(-1, b.size(-2), b.size(-1))
        rep = enc(x, padding_mask=pm, attn_mask=b)[0]
        for h_ in hooks:
            h_.remove()
        g = torch.randn(rep.shape, generator=torch.Generator().manual_seed(31 + seed))
        (rep * g).sum().backward()
        named = dict(mods.named_parameters())
        gsel = ["embed_tokens.weight", "gbf.means.weight", "gbf.stds.weight", "gbf.mul.weight", "gbf.bias.weight",
                "gbf_proj.linear1.weight", "gbf_proj.linear2.weight",
                "encoder.layers.0.self_attn.in_proj.weight", "encoder.layers.0.self_attn.in_proj.bias",
                "encoder.layers.14.self_attn.in_proj.weight", "encoder.layers.14.self_attn.in_proj.bias",
                "encoder.layers.7.fc1.weight", "encoder.layers.7.fc2.bias", "encoder.layers.0.self_attn.out_proj.weight",
                "encoder.layers.3.self_attn_layer_norm.weight", "encoder.layers.14.final_layer_norm.bias",
                "encoder.emb_layer_norm.weight", "encoder.final_layer_norm.weight"]
        d = {"in.tokens": tokens, "in.dist": dist, "in.edge_type": et, "in.up": g, "out.rep": rep,
             "out.x_layers_mol0": torch.stack(per_layer_x), "out.pair_layers_mol0_head0": torch.stack(per_layer_pair),
             "cfg": np.array([H, D, Fd, nl, seed, ROWS]), "cfg.std": np.array([std])}
        for k in gsel:
            gr = named[k].grad
            d["grad." + k] = gr[:ROWS] if (gr.dim() == 2 and gr.shape[0] > ROWS and gr.numel() > 70000) else gr
        # pin the restatement at this depth
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        mine = restate.unimol_encoder(tokens, dist, et, p, heads=H, n_layers=nl)
        _close(mine, rep, 1e-5, "enc15.%s.rep" % tag)
        (mine * g).sum().backward()
        for k in gsel:
            _close(p[k].grad, named[k].grad, 1e-4, "enc15.%s.grad.%s" % (tag, k))
        # float64 truth (the restatement, pinned above, evaluated in double): at this depth the fp32 reference's own
        # round-off (cancellation in the gbf.* gradients: sums of signed terms over 17 k pairs x 64 heads) is what limits
        # an fp32-vs-fp32 comparison; the test bounds |ours - truth64| by the reference's own |ref32 - truth64|
        p64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
        rep64 = restate.unimol_encoder(tokens, dist.double(), et, p64, heads=H, n_layers=nl, wide=True)
        (rep64 * g.double()).sum().backward()
        d["out.rep64"] = rep64.detach().float()
        for k in gsel:
            gr = p64[k].grad
            gr = gr[:ROWS] if (gr.dim() == 2 and gr.shape[0] > ROWS and gr.numel() > 70000) else gr
            d["grad64." + k] = gr.float()
            e32 = ((d["grad." + k].double() - gr).norm() / gr.norm()).item()
            print("   %-50s ref32 vs truth64: %.1e" % (k, e32))
        _save("encoder_15L_" + tag, d)


def gold_infonce(ref):
    inf = ref["infonce"]
    g = torch.Generator().manual_seed(40)
    for tag, N, De in (("n16", 16, 50), ("n37", 37, 50), ("n64d512", 64, 512)):
        q = torch.randn(N, De, generator=g).requires_grad_(True)
        k = (0.5 * q.detach() + torch.randn(N, De, generator=g)).requires_grad_(True)
        loss = inf.info_nce(q, k, temperature=0.1)
        loss.backward()
        q2 = q.detach().clone().requires_grad_(True)
        k2 = k.detach().clone().requires_grad_(True)
        mine = restate.info_nce(q2, k2, 0.1)
        mine.backward()
        _close(mine, loss, 1e-6, "info_nce")
        _close(q2.grad, q.grad, 1e-6, "info_nce.dq")
        _save("infonce_" + tag, {"in.q": q, "in.k": k, "out.loss": loss, "grad.q": q.grad, "grad.k": k.grad})
    # the module: 512 -> 512 -> gelu -> 50, unmasked mean over the sequence (Q3)
    torch.manual_seed(41)
    mod = inf.InfoNCE(512, 512)
    mod.eval()
    B, L, S = 6, 7, 9
    query = torch.randn(B, L, 512, generator=g).requires_grad_(True)
    pos = torch.randn(B, S, 512, generator=g).requires_grad_(True)
    loss = mod(query, pos)
    loss.backward()
    sd = {"infonce." + k_: v for k_, v in mod.state_dict().items()}
    p = {k_: v.clone() for k_, v in sd.items()}
    mine = restate.infonce_head(query.detach(), pos.detach(), p)
    _close(mine, loss, 1e-6, "InfoNCE.forward")
    d = {"in.query": query, "in.positive": pos, "out.loss": loss, "grad.query": query.grad,
         "grad.positive": pos.grad}
    d.update({"w." + k_: (v.half() if v.numel() > 100000 else v) for k_, v in sd.items()})
    _save("infonce_module", d) if False else None     # weights too big to be useful; skip
    # error behaviour (models/infonce.py:45-67)
    for bad in ((torch.randn(4), torch.randn(4, 3)), (torch.randn(4, 3), torch.randn(5, 3)),
                (torch.randn(4, 3), torch.randn(4, 2))):
        for fn in (inf.info_nce, restate.info_nce):
            try:
                fn(*bad)
                raise AssertionError("expected ValueError")
            except ValueError:
                pass


def gold_ct(ref):
    ct = ref["contrastive"]
    g = torch.Generator().manual_seed(50)
    for N in (16, 45):
        f = torch.randn(N, 512, generator=g)
        y = torch.randn(N, 1, generator=g)
        yhat = y + 0.3 * torch.randn(N, 1, generator=g)
        wts = torch.rand(N, generator=g) + 0.5
        wts = wts / wts.mean()
        cls = torch.randint(0, 3, (N, 1), generator=g)
        multi = torch.randint(0, 2, (N, 5), generator=g).float()
        cases = {
            "regress_w": (ct.CT_Regress, restate.ct_regress, (y, yhat), dict(weights=wts, w=0.2)),
            "regress_now": (ct.CT_Regress, restate.ct_regress, (y, yhat), dict(w=0.2)),
            "single": (ct.CT_Single, restate.ct_single, (cls, yhat), dict()),
            "multi": (ct.CT_Multi, restate.ct_multi, (multi, yhat), dict()),
        }
        d = {"in.feature": f, "in.y": y, "in.yhat": yhat, "in.weights": wts, "in.cls": cls, "in.multi": multi}
        for name, (rf, mf, args, kw) in cases.items():
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                f1 = f.clone().requires_grad_(True)
                loss = rf(f1, args[0], args[1], **kw)
                loss.backward()
            f2 = f.clone().requires_grad_(True)
            kw2 = {k_: v for k_, v in kw.items()}
            mine = mf(f2, args[0], args[1], **kw2) if name.startswith("regress") else mf(f2, args[0], None, **kw2)
            mine.backward()
            _close(mine, loss, 1e-6, "ct." + name)
            _close(f2.grad, f1.grad, 1e-6, "ct.grad." + name)
            d["out." + name] = loss
            d["grad." + name] = f1.grad
        # masks, bit-exact
        pos, neg, _ = restate.ct_masks("regress", y, yhat, 0.2)
        l_dist = (y - y.T).abs()
        p_dist = (yhat - yhat.T).abs()
        rpos = l_dist.le(0.2)
        rneg = (~l_dist.le(0.2)) * p_dist.le(0.2)
        for i in range(N):
            rpos[i][i] = 0
        assert torch.equal(pos, rpos) and torch.equal(neg, rneg)
        d["out.regress_pos"], d["out.regress_neg"] = rpos, rneg
        spos, sneg, _ = restate.ct_masks("single", cls)
        d["out.single_pos"], d["out.single_neg"] = spos, sneg
        mpos, mneg, _ = restate.ct_masks("multi", multi)
        d["out.multi_pos"], d["out.multi_neg"] = mpos, mneg
        _save("ct_n%d" % N, d)


def gold_fds(ref):
    fds_mod, util = ref["fds"], ref["util"]
    g = torch.Generator().manual_seed(60)
    raw = torch.randn(400, generator=g).double() * 1.7 + 0.3
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "train.csv")
        with open(csv, "w") as fh:
            fh.write("expt\n" + "\n".join("%.17g" % v for v in raw.tolist()) + "\n")
        orig_to = torch.Tensor.to

        def to_nocuda(self, *a, **k):       # models/fds.py:84 hard-codes .to('cuda')
            if a and a[0] == "cuda":
                return self
            return orig_to(self, *a, **k)

        torch.Tensor.to = to_nocuda
        try:
            F_ = fds_mod.FDS(feature_dim=16, raw_data=csv, col_data="expt", using_scale=True,
                             bucket_num=12, bucket_start=0, start_update=0, start_smooth=1,
                             kernel="gaussian", ks=5, sigma=1, momentum=0.9)
        finally:
            torch.Tensor.to = orig_to
    cfg = dict(min_value=float(F_.min_value), bin_width=float(F_.bin_width), bucket_num=12,
               bucket_start=0, start_update=0, start_smooth=1, momentum=0.9)
    win = restate.fds_kernel_window("gaussian", 5, 1)
    _close(win, F_.kernel_window, 1e-6, "fds.window")
    st = {k: v.clone() for k, v in F_.state_dict().items()}
    d = {"cfg.min_value": cfg["min_value"], "cfg.bin_width": cfg["bin_width"], "cfg.window": F_.kernel_window}
    # labels stress: exact bin edges, far tails (edge clamp, Q10), interior
    n = 96
    labels = torch.randn(n, 1, generator=g) * 1.2
    edges = torch.tensor([cfg["min_value"] + k_ * cfg["bin_width"] for k_ in range(0, 13)], dtype=torch.float64)
    labels[:13, 0] = edges.float()
    labels[13, 0] = -9.0
    labels[14, 0] = 9.0
    feats_e0 = torch.randn(n, 16, generator=g)
    feats_e0[:, 3] = 0.25                      # a zero-variance column (Q13 middle branch)
    feats_e1 = torch.randn(n, 16, generator=g) * 1.5 + 0.2
    bins_ref = torch.Tensor([int((v - F_.min_value) // F_.bin_width) for v in labels[:, 0]])
    bins = restate.fds_label_bins(labels, cfg["min_value"], cfg["bin_width"])
    assert torch.equal(bins, bins_ref.long()), "bin restatement differs"
    d["in.labels"], d["in.feats_e0"], d["in.feats_e1"], d["out.bins"] = labels, feats_e0, feats_e1, bins
    # epoch 0: stats pass;  epoch 1: last-epoch update, smooth, stats pass;  epoch 2: again
    F_.update_running_stats(feats_e0, labels, 0)
    restate.fds_update_running_stats(feats_e0, labels, 0, st, cfg)
    F_.update_last_epoch_stats(1)
    restate.fds_update_last_epoch_stats(1, st, win)
    for k_ in ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked", "epoch"):
        _close(st[k_], getattr(F_, k_), 1e-6, "fds.e0." + k_)
        d["out.e1." + k_] = getattr(F_, k_).clone()
    x = torch.randn(n, 16, generator=g).requires_grad_(True)
    xs = F_.smooth(x * 1.0, labels, 1)
    up = torch.randn(n, 16, generator=g)
    (xs * up).sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    xs2 = restate.fds_smooth(x2 * 1.0, labels, 1, st, cfg)
    (xs2 * up).sum().backward()
    _close(xs2, xs, 1e-6, "fds.smooth")
    _close(x2.grad, x.grad, 1e-6, "fds.smooth.grad")
    d["in.smooth_x"], d["in.smooth_up"], d["out.smooth"], d["grad.smooth_x"] = x, up, xs, x.grad
    # batch WITHOUT the edge bins: tails must be left untouched (Q10)
    sub = (bins > 0) & (bins < 11) | (labels[:, 0].abs() > 8)
    xs3 = F_.smooth(x.detach()[sub].clone(), labels[sub], 1)
    xs4 = restate.fds_smooth(x.detach()[sub].clone(), labels[sub], 1, st, cfg)
    _close(xs4, xs3, 1e-6, "fds.smooth.noedge")
    d["in.sub"], d["out.smooth_noedge"] = sub, xs3
    F_.update_running_stats(feats_e1, labels, 1)
    restate.fds_update_running_stats(feats_e1, labels, 1, st, cfg)
    F_.update_last_epoch_stats(2)
    restate.fds_update_last_epoch_stats(2, st, win)
    for k_ in ("running_mean", "running_var", "smoothed_mean_last_epoch", "smoothed_var_last_epoch", "num_samples_tracked", "epoch"):
        _close(st[k_], getattr(F_, k_), 1e-6, "fds.e1." + k_)
        d["out.e2." + k_] = getattr(F_, k_).clone()
    assert F_.running_mean_last_epoch is F_.running_mean        # Q9 aliasing
    # calibrate_mean_var branches
    m1, v1, m2, v2 = torch.randn(16, generator=g), torch.rand(16, generator=g), torch.randn(16, generator=g), torch.rand(16, generator=g) * 30
    mat = torch.randn(5, 16, generator=g)
    d["in.cal_mat"], d["in.cal_m1"], d["in.cal_v1"], d["in.cal_m2"], d["in.cal_v2"] = mat, m1, v1, m2, v2
    d["out.cal_full"] = util.calibrate_mean_var(mat.clone(), m1, v1, m2, v2)
    v1z = v1.clone()
    v1z[[2, 7]] = 0
    d["out.cal_zero_cols"] = util.calibrate_mean_var(mat.clone(), m1, v1z, m2, v2)
    d["out.cal_tiny"] = util.calibrate_mean_var(mat.clone(), m1, v1 * 0, m2, v2)
    _close(restate.calibrate_mean_var(mat.clone(), m1, v1, m2, v2), d["out.cal_full"], 1e-7, "cal.full")
    _close(restate.calibrate_mean_var(mat.clone(), m1, v1z, m2, v2), d["out.cal_zero_cols"], 1e-7, "cal.zero")
    _close(restate.calibrate_mean_var(mat.clone(), m1, v1 * 0, m2, v2), d["out.cal_tiny"], 1e-7, "cal.tiny")
    _save("fds", d)


def gold_featurise(ref):
    """data/conformer.py coords2unimol executed from the reference tree (rdkit / config / logger stubbed: the function
    itself only needs numpy + scipy + the dictionary), batched with the reference's own padding helpers."""
    import importlib.util
    import types
    stubs = {}
    for name in ("rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.RDLogger", "config"):
        if name not in sys.modules:
            stubs[name] = sys.modules[name] = types.ModuleType(name)
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["rdkit"].RDLogger = sys.modules["rdkit.RDLogger"]
    sys.modules["rdkit.Chem"].AllChem = sys.modules["rdkit.Chem.AllChem"]
    sys.modules["rdkit.RDLogger"].DisableLog = lambda *a, **k: None
    sys.modules["config"].MODEL_CONFIG = {}
    spec = importlib.util.spec_from_file_location("_ref_conformer", os.path.join(ref_loader.REF_ROOT, "data", "conformer.py"))
    mod = importlib.util.module_from_spec(spec)
    with ref_loader._tmp_cwd():
        spec.loader.exec_module(mod)
    for name in stubs:
        sys.modules.pop(name, None)
    from unicore.data import Dictionary
    symbols = ["[PAD]", "[CLS]", "[SEP]", "[UNK]", "C", "N", "O", "S", "H", "Cl", "F", "Br", "I", "Si", "P", "B", "Na", "K",
               "Al", "Ca", "Sn", "As", "Hg", "Fe", "Zn", "Cr", "Se", "Gd", "Au", "Li"]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "mol.dict.txt")
        open(path, "w").write("\n".join(symbols) + "\n")
        dictionary = Dictionary.load(path)
    dictionary.add_symbol("[MASK]", is_special=True)
    assert len(dictionary) == 31
    rng = np.random.RandomState(11)
    samples = []
    for n in (5, 12, 30, 64, 3, 20, 1):
        atoms = [symbols[i] for i in rng.randint(4, 30, size=n)]
        atoms = [a if a != "H" else "C" for a in atoms]            # remove_hs would drop hydrogens
        xyz = (rng.randn(n, 3) * 2.0 + rng.randn(1, 3) * 5.0).astype(np.float64)
        samples.append(mod.coords2unimol(atoms, xyz, dictionary, max_atoms=256, remove_hs=True))
    util = ref["util"]
    tok = util.pad_1d_tokens([torch.as_tensor(s["src_tokens"]).long() for s in samples], pad_idx=0)
    dist = util.pad_2d([torch.as_tensor(s["src_distance"]).float() for s in samples], pad_idx=0)
    et = util.pad_2d([torch.as_tensor(s["src_edge_type"]).long() for s in samples], pad_idx=0)
    coord = util.pad_coords([torch.as_tensor(s["src_coord"]).float() for s in samples], pad_idx=0.0)
    d2, e2 = restate.featurise(tok, coord, n_dict=len(dictionary), pad_idx=0)
    assert torch.equal(d2, dist), "featurise: distances differ from the reference (max %g)" % (d2 - dist).abs().max()
    assert torch.equal(e2, et), "featurise: edge types differ from the reference"
    _save("featurise", {"in.src_tokens": tok, "in.src_coord": coord, "out.src_distance": dist, "out.src_edge_type": et})


def gold_loss(ref):
    """Task losses (models/loss.py:9-289) and the explicit-negatives / reduction='none' branches of info_nce
    (models/infonce.py:71-88): outputs of the reference's own functions on seeded inputs; the host-side mirrors in
    mmdti_b200/models are checked against them on CPU (tests/test_host_logic.py)."""
    L, inf = ref["loss"], ref["infonce"]
    g = torch.Generator().manual_seed(70)
    B, C = 24, 5
    x = torch.randn(B, C, generator=g)
    y = torch.randn(B, C, generator=g)
    yb = (torch.rand(B, C, generator=g) < 0.3).float()
    yb_nan = yb.clone()
    yb_nan[torch.rand(B, C, generator=g) < 0.2] = float("nan")
    y_nan = y.clone()
    y_nan[torch.rand(B, C, generator=g) < 0.2] = float("nan")
    prob = torch.rand(B, generator=g)
    cls = torch.randint(0, C, (B, 1), generator=g)
    d = {"in.x": x, "in.y": y, "in.yb": yb, "in.yb_nan": yb_nan, "in.y_nan": y_nan, "in.prob": prob, "in.cls": cls}
    d["out.rmse"] = L.RMSELoss()(x, y)
    ghmc, ghmr = L.GHMC_Loss(bins=10, alpha=0.5), L.GHMR_Loss(bins=10, alpha=0.5, mu=0.02)
    d["out.ghmc_1"], d["out.ghmc_2"] = ghmc(x, yb), ghmc(0.5 * x, yb)       # second call exercises the EMA of the bin counts
    d["out.ghmr_1"], d["out.ghmr_2"] = ghmr(x, y), ghmr(0.5 * x, y)
    d["out.masked_bce"] = L.MaskedBCEWithLogitsLoss()(x, yb_nan)
    d["out.mae_nan"] = L.MAEwithNan(x, y_nan)
    d["out.bce_nan"] = L.BCEwithNan(x, yb_nan)
    d["out.focal"] = L.FocalLoss(prob, yb[:, 0])
    d["out.focal_logits"] = L.FocalLossWithLogits(x, yb_nan)
    d["out.ce"] = L.myCrossEntropyLoss(x, cls)
    # the symmetric form (infonce.py:98) transposes the (N, 1+M) logits against N labels: the reference's explicit-negatives
    # branches only run when M + 1 == N (anything else raises inside F.cross_entropy)
    N, De, M = 12, 20, 11
    q, k = torch.randn(N, De, generator=g), torch.randn(N, De, generator=g)
    neg_u, neg_p = torch.randn(M, De, generator=g), torch.randn(N, M, De, generator=g)
    d.update({"in.q": q, "in.k": k, "in.neg_unpaired": neg_u, "in.neg_paired": neg_p})
    d["out.nce_unpaired"] = inf.info_nce(q, k, neg_u, temperature=0.1, negative_mode="unpaired")
    d["out.nce_paired"] = inf.info_nce(q, k, neg_p, temperature=0.2, negative_mode="paired")
    d["out.nce_none"] = inf.info_nce(q, k, temperature=0.1, reduction="none")
    _save("loss", d)


def gold_cross_modal(ref):
    """the reference's CrossAttentionModel (models/mm_model.py:379-406; BertCrossEncoder,
    models/mm_module.py:663-677) at its own configuration (crossmodal_config: hidden 512, 16 heads, FFN 2048, eps 1e-12),
    eval mode (dropout off), ragged masks, followed by the masked pooling of models/mm_model.py:572-576.
    Weights: oracle/detw.py, std 0.05 (6.3 M parameters, not stored).  Stored: both outputs, the pooled features, the
    gradients of both inputs and 12 parameter gradients (rows [:ROWS] of the matrices)."""
    from oracle.detw import det_state_dict, det_tensor
    mm = ref["mm_model"]
    cfg = mm.crossmodal_config()
    ROWS = 32
    net = mm.CrossAttentionModel(cfg, num_layers=1)
    sd = det_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=21, std=0.05)
    net.load_state_dict(sd)
    net.eval()
    B, L1, L2, D = 3, 13, 10, cfg.hidden_size
    x1 = det_tensor((B, L1, D), 901, std=1.0).requires_grad_(True)          # "text_embeddings" slot (the graph tokens in MM_Model)
    x2 = det_tensor((B, L2, D), 902, std=1.0).requires_grad_(True)
    m1 = torch.zeros(B, L1, dtype=torch.bool)
    m2 = torch.zeros(B, L2, dtype=torch.bool)
    for b, (n1, n2) in enumerate(((13, 4), (7, 10), (1, 6))):
        m1[b, :n1] = True
        m2[b, :n2] = True
    t2g, g2t = net(x1, x2, m1, m2)
    a, c = t2g.clone(), g2t.clone()
    a[~m1] = 0.0
    c[~m2] = 0.0
    final = torch.cat((a, c), dim=1)
    pooled = final.sum(dim=1) / (m1.sum(dim=1).view(-1, 1) + m2.sum(dim=1).view(-1, 1))
    up = det_tensor((B, D), 903, std=1.0)
    (pooled * up).sum().backward()
    named = dict(net.named_parameters())
    gsel = ["text_attention.layer.0.attention.self.query.weight", "text_attention.layer.0.attention.self.key.bias",
            "text_attention.layer.0.attention.self.value.weight", "text_attention.layer.0.attention.output.dense.weight",
            "text_attention.layer.0.attention.output.LayerNorm.weight", "text_attention.layer.0.intermediate.dense.bias",
            "text_attention.layer.0.output.dense.weight", "text_attention.layer.0.output.LayerNorm.bias",
            "graph_attention.layer.0.attention.self.query.bias", "graph_attention.layer.0.attention.self.key.weight",
            "graph_attention.layer.0.intermediate.dense.weight", "graph_attention.layer.0.output.LayerNorm.weight"]
    d = {"in.x1": x1, "in.x2": x2, "in.m1": m1, "in.m2": m2, "in.up": up, "out.t2g": t2g, "out.g2t": g2t, "out.pooled": pooled,
         "grad.x1": x1.grad, "grad.x2": x2.grad, "cfg": np.array([cfg.num_attention_heads, D, cfg.intermediate_size, 21, ROWS])}
    for k in gsel:
        d["grad." + k] = named[k].grad[:ROWS] if named[k].grad.dim() == 2 else named[k].grad
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y1, y2 = x1.detach().clone().requires_grad_(True), x2.detach().clone().requires_grad_(True)
    mt2g, mg2t = restate.cross_modal(y1, y2, m1, m2, p, heads=cfg.num_attention_heads, eps=cfg.layer_norm_eps)
    _close(mt2g, t2g, 2e-6, "cross.t2g")
    _close(mg2t, g2t, 2e-6, "cross.g2t")
    mp = restate.fuse_pool(mt2g, mg2t, m1, m2)
    _close(mp, pooled, 2e-6, "cross.pooled")
    (mp * up).sum().backward()
    _close(y1.grad, x1.grad, 2e-5, "cross.grad.x1")
    _close(y2.grad, x2.grad, 2e-5, "cross.grad.x2")
    for k in gsel:
        _close(p[k].grad, named[k].grad, 2e-5, "cross.grad." + k)
    _save("cross_modal", d)


def gold_chemberta(ref):
    """the second-modality encoder as the reference runs it -- Hugging Face ``RobertaModel``
    (models/mm_model.py:475,562), here a random-init 512-d stand-in (2 layers, 8 heads x 64, FFN 1024, vocab 64; the
    reference's heads expect a 512-d ChemBERTa, models/mm_model.py:493), eval mode, ragged padding.
    Weights: oracle/detw.py std 0.05 (not stored).  Stored: last_hidden_state, the gradient of the word / position
    embeddings and 6 layer-parameter gradients (rows [:ROWS])."""
    from transformers import RobertaConfig, RobertaModel
    from oracle.detw import det_state_dict, det_tensor
    ROWS, H, D, Fd, nl, V, P = 32, 8, 512, 1024, 2, 64, 48
    cfg = RobertaConfig(vocab_size=V, hidden_size=D, num_hidden_layers=nl, num_attention_heads=H, intermediate_size=Fd,
                        max_position_embeddings=P, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pad_token_id=1)
    net = RobertaModel(cfg)
    sd = det_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=31, std=0.05)
    net.load_state_dict(sd)
    net.eval()
    B, S = 3, 21
    g = torch.Generator().manual_seed(77)
    ids = torch.randint(4, V, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    for b, n in enumerate((21, 9, 2)):
        am[b, n:] = 0
        ids[b, n:] = 1
    out = net(ids, am, return_dict=True)[0]
    up = det_tensor((B, S, D), 904, std=1.0) * am[..., None]
    (out * up).sum().backward()
    named = dict(net.named_parameters())
    gsel = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight", "embeddings.LayerNorm.weight",
            "encoder.layer.0.attention.self.query.weight", "encoder.layer.0.attention.self.value.bias",
            "encoder.layer.0.output.dense.weight", "encoder.layer.1.attention.output.dense.weight",
            "encoder.layer.1.intermediate.dense.bias", "encoder.layer.1.output.LayerNorm.weight"]
    d = {"in.ids": ids, "in.mask": am, "in.up": up, "out.hidden": out, "cfg": np.array([H, D, Fd, nl, V, P, 31, ROWS])}
    for k in gsel:
        gr = named[k].grad
        d["grad." + k] = gr[:ROWS] if (gr.dim() == 2 and "embeddings" not in k) else gr
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    mine = restate.roberta_encoder(ids, am, p, heads=H, n_layers=nl, eps=cfg.layer_norm_eps, pad_idx=1)
    _close(mine * am[..., None], out * am[..., None], 2e-6, "chemberta.hidden")
    (mine * up).sum().backward()
    for k in gsel:
        _close(p[k].grad, named[k].grad, 2e-5, "chemberta.grad." + k)
    _save("chemberta", d)


def main():
    torch.set_num_threads(8)
    ref = ref_loader.load()
    gold_pair_bias(ref)
    gold_encoder(ref)
    gold_encoder_slice(ref)
    gold_encoder_15l(ref)
    gold_infonce(ref)
    gold_ct(ref)
    gold_fds(ref)
    gold_featurise(ref)
    gold_loss(ref)
    gold_cross_modal(ref)
    gold_chemberta(ref)
    print("all fixtures written and the restatement reproduces each of them")


if __name__ == "__main__":
    main()
