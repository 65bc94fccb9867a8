"""CPU restatement of the MM-DTI training hot path (TEST INFRASTRUCTURE — see
oracle/__init__.py).  Plain PyTorch on the host, functional style, parameters passed
as dicts keyed by the reference's ``state_dict`` names so weights can be shared with
the reference modules and with the CUDA drop-ins.  Gradients come from autograd.

Every function cites the reference lines it follows (paths relative to /root/reference).
Uni-Core pieces (attention layer, LayerNorm) follow the public algorithm restated in
oracle/shims/unicore/modules.py (third-party, unpinned).
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- K1
GAUSS_PI = 3.14159  # truncated on purpose: models/mm_model.py:222


def featurise(src_tokens, src_coord, n_dict=31, pad_idx=0):
    """Pair features of a PADDED batch from tokens (B,L) and centred coordinates (B,L,3) f32:
    src_distance (B,L,L) f32 and src_edge_type (B,L,L) int64, zero wherever either position is padding.
    Follows data/conformer.py:205-212,216-218 (scipy distance_matrix = float64 (sum |d|^2)^(1/2) of the
    float32-valued coordinates, cast to float32; edge type = tok_i * len(dictionary) + tok_j) and the
    zero padding of utils/util.py:41-105 / models/mm_model.py:656-661."""
    import numpy as np
    tok = np.asarray(src_tokens, dtype=np.int64)
    xyz = np.asarray(src_coord, dtype=np.float32).astype(np.float64)
    d = xyz[:, :, None, :] - xyz[:, None, :, :]
    sq = np.abs(d) ** 2
    dist = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])
    valid = (tok != pad_idx)
    pair = valid[:, :, None] & valid[:, None, :]
    et = tok[:, :, None] * n_dict + tok[:, None, :]
    return torch.from_numpy(np.where(pair, dist, 0.0).astype(np.float32)), torch.from_numpy(np.where(pair, et, 0))


def gaussian(x, mean, std):
    """models/mm_model.py:211-224."""
    a = (2 * GAUSS_PI) ** 0.5
    return torch.exp(-0.5 * (((x - mean) / std) ** 2)) / (a * std)


def gaussian_basis(dist, edge_type, p, prefix="gbf.", wide=False):
    """GaussianLayer.forward, models/mm_model.py:254-269.
    dist (B,L,L) float, edge_type (B,L,L) int64 -> (B,L,L,K) fp32.
    wide=True: the same expressions with the reference's ``.float()`` casts replaced by float64 (dist and the
    parameters given as float64) -- a higher-precision truth for sums over > 1e5 pairs; not the reference's arithmetic."""
    cast = torch.float64 if wide else torch.float32
    mul = p[prefix + "mul.weight"][edge_type].to(dist.dtype)      # (B,L,L,1)
    bias = p[prefix + "bias.weight"][edge_type].to(dist.dtype)
    u = mul * dist.unsqueeze(-1) + bias
    K = p[prefix + "means.weight"].shape[-1]
    u = u.expand(-1, -1, -1, K)
    mean = p[prefix + "means.weight"].to(cast).view(-1)
    std = p[prefix + "stds.weight"].to(cast).view(-1).abs() + 1e-5
    return gaussian(u.to(cast), mean, std).to(p[prefix + "means.weight"].dtype)


def nonlinear_head(x, p, prefix="gbf_proj."):
    """NonLinearHead.forward (gelu), models/mm_model.py:117-128."""
    x = F.linear(x, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"])
    x = F.gelu(x)
    return F.linear(x, p[prefix + "linear2.weight"], p[prefix + "linear2.bias"])


def pair_bias(dist, edge_type, p, wide=False):
    """models/mm_model.py:553-556: gbf -> gbf_proj -> permute(0,3,1,2) -> (B*H,L,L)."""
    g = gaussian_basis(dist, edge_type, p, wide=wide)
    o = nonlinear_head(g, p)
    o = o.permute(0, 3, 1, 2).contiguous()
    return o.view(-1, o.size(-2), o.size(-1))


# --------------------------------------------------------------------------- K2
def _drop(x, p, mask):
    """dropout with an explicit keep-mask (for replaying the CUDA Philox mask)."""
    if mask is None or p == 0.0:
        return x
    return x * mask.to(x.dtype) / (1.0 - p)


def pair_attention(q, k, v, bias, scale, attn_dropout=0.0, keep=None):
    """Uni-Core SelfMultiheadAttention core with return_attn=True
    (oracle/shims/unicore/modules.py; call site models/transformers.py:137-139).
    q,k,v (B,H,L,d); bias (B,H,L,L) (−inf at padded keys) ->
    (o (B,H,L,d), scores (B,H,L,L)).  scores = (q*scale)·kᵀ + bias is what the layer
    RETURNS and what becomes the next layer's bias."""
    s = torch.matmul(q * scale, k.transpose(-1, -2)) + bias
    a = torch.softmax(s, dim=-1)
    a = _drop(a, attn_dropout, keep)
    return torch.matmul(a, v), s


def encoder_layer(x, bias, p, prefix, heads, attn_dropout=0.0, dropout=0.0, keeps=None):
    """Uni-Core TransformerEncoderLayer, pre-LN, return_attn=True.
    x (B,L,D); bias (B*H,L,L) -> (x', scores (B*H,L,L)).
    keeps: optional dict of keep-masks {'attn','res1','res2'}."""
    keeps = keeps or {}
    B, L, D = x.shape
    d = D // heads
    r = x
    h = F.layer_norm(x, (D,), p[prefix + "self_attn_layer_norm.weight"],
                     p[prefix + "self_attn_layer_norm.bias"], 1e-5)
    qkv = F.linear(h, p[prefix + "self_attn.in_proj.weight"], p[prefix + "self_attn.in_proj.bias"])
    q, k, v = qkv.chunk(3, dim=-1)

    def split(t):
        return t.view(B, L, heads, d).transpose(1, 2)

    o, s = pair_attention(split(q), split(k), split(v), bias.view(B, heads, L, L), d ** -0.5,
                          attn_dropout, keeps.get("attn"))
    o = o.transpose(1, 2).reshape(B, L, D)
    o = F.linear(o, p[prefix + "self_attn.out_proj.weight"], p[prefix + "self_attn.out_proj.bias"])
    x = r + _drop(o, dropout, keeps.get("res1"))
    r = x
    h = F.layer_norm(x, (D,), p[prefix + "final_layer_norm.weight"],
                     p[prefix + "final_layer_norm.bias"], 1e-5)
    h = F.gelu(F.linear(h, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    h = F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])
    x = r + _drop(h, dropout, keeps.get("res2"))
    return x, s.reshape(B * heads, L, L)


def merge_key_padding(bias, padding_mask, heads, fill=float("-inf")):
    """models/transformers.py:122-132: in-place −inf fill of padded KEY columns."""
    if padding_mask is None:
        return bias
    B, L = padding_mask.shape
    bias.view(B, heads, L, L).masked_fill_(padding_mask[:, None, None, :].to(torch.bool), fill)
    return bias


def _norm_loss(x, eps=1e-10, tolerance=1.0):
    """models/transformers.py:141-146."""
    x = x.float()
    max_norm = x.shape[-1] ** 0.5
    norm = torch.sqrt(torch.sum(x ** 2, dim=-1) + eps)
    return F.relu((norm - max_norm).abs() - tolerance)


def _masked_mean(mask, value, dim=-1, eps=1e-10):
    """models/transformers.py:148-151."""
    return (torch.sum(mask * value, dim=dim) / (eps + torch.sum(mask, dim=dim))).mean()


def encoder_with_pair(emb, attn_mask, padding_mask, p, heads, n_layers, prefix="encoder.",
                      final_head_ln=False, keeps=None, emb_dropout=0.0, attn_dropout=0.0,
                      dropout=0.0):
    """TransformerEncoderWithPair.forward, models/transformers.py:96-183.
    Mutates ``attn_mask`` in place exactly like the reference (Q1).
    Returns (x, pair (B,L,L,H), delta_pair (B,L,L,H), x_norm, delta_pair_norm)."""
    keeps = keeps or {}
    B, L, D = emb.shape
    x = F.layer_norm(emb, (D,), p[prefix + "emb_layer_norm.weight"],
                     p[prefix + "emb_layer_norm.bias"], 1e-5)
    x = _drop(x, emb_dropout, keeps.get("emb"))
    if padding_mask is not None:
        x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))
    input_attn_mask = attn_mask
    assert attn_mask is not None
    attn_mask = merge_key_padding(attn_mask, padding_mask, heads)
    for i in range(n_layers):
        x, attn_mask = encoder_layer(x, attn_mask, p, "%slayers.%d." % (prefix, i), heads,
                                     attn_dropout, dropout, keeps.get(i))
    x_norm = _norm_loss(x)
    if padding_mask is not None:
        token_mask = 1.0 - padding_mask.float()
    else:
        token_mask = torch.ones_like(x_norm)
    x_norm = _masked_mean(token_mask, x_norm)
    if (prefix + "final_layer_norm.weight") in p:
        x = F.layer_norm(x, (D,), p[prefix + "final_layer_norm.weight"],
                         p[prefix + "final_layer_norm.bias"], 1e-5)
    delta = attn_mask - input_attn_mask          # NaN at −inf columns …
    delta = merge_key_padding(delta, padding_mask, heads, 0)   # … then zeroed
    pair = attn_mask.view(B, -1, L, L).permute(0, 2, 3, 1).contiguous()
    delta = delta.view(B, -1, L, L).permute(0, 2, 3, 1).contiguous()
    pair_mask = token_mask[..., None] * token_mask[..., None, :]
    delta_norm = _masked_mean(pair_mask, _norm_loss(delta), dim=(-1, -2))
    if final_head_ln:
        delta = F.layer_norm(delta, (heads,), p[prefix + "final_head_layer_norm.weight"],
                             p[prefix + "final_head_layer_norm.bias"], 1e-5)
    return x, pair, delta, x_norm, delta_norm


def unimol_encoder(src_tokens, src_distance, src_edge_type, p, heads=64, n_layers=15, pad_idx=0, wide=False):
    """UnimolEncoder.forward (models/encoder.py:458-502) == models/mm_model.py:545-559:
    padding mask -> embed -> pair bias -> encoder -> all_repr (B,L,D).
    wide=True (float64 parameters and distances): the same expressions evaluated in float64 throughout, used as a
    higher-precision truth when the fp32 reference's own round-off is what limits a comparison."""
    padding_mask = src_tokens.eq(pad_idx)
    if not padding_mask.any():
        padding_mask = None
    x = F.embedding(src_tokens, p["embed_tokens.weight"], padding_idx=pad_idx)
    bias = pair_bias(src_distance, src_edge_type, p, wide=wide)
    return encoder_with_pair(x, bias, padding_mask, p, heads, n_layers)[0]


# --------------------------------------------------------------------------- K3
def info_nce(query, positive_key, temperature=0.1, reduction="mean"):
    """models/infonce.py:42-98 with negative_keys=None (the only mode MM_Model uses)."""
    if query.dim() != 2:
        raise ValueError("<query> must have 2 dimensions.")
    if positive_key.dim() != 2:
        raise ValueError("<positive_key> must have 2 dimensions.")
    if len(query) != len(positive_key):
        raise ValueError("<query> and <positive_key> must must have the same number of samples.")
    if query.shape[-1] != positive_key.shape[-1]:
        raise ValueError("Vectors of <query> and <positive_key> should have the same number of components.")
    q = F.normalize(query, dim=-1)
    k = F.normalize(positive_key, dim=-1)
    logits = q @ k.transpose(-2, -1)
    labels = torch.arange(len(q), device=q.device)
    return (F.cross_entropy(logits / temperature, labels, reduction=reduction)
            + F.cross_entropy(logits.T / temperature, labels, reduction=reduction)) / 2


def infonce_head(query, positive, p, prefix="infonce.", keep=None, embed_dropout=0.0,
                 temperature=0.1):
    """InfoNCE.forward, models/infonce.py:23-38: dropout(query) -> two 2-layer GELU MLPs
    -> UNMASKED mean over the sequence axis -> info_nce."""
    x = _drop(query, embed_dropout, keep)

    def mlp(t, pre):
        t = F.gelu(F.linear(t, p[pre + "0.weight"], p[pre + "0.bias"]))
        return F.linear(t, p[pre + "2.weight"], p[pre + "2.bias"])

    pq = mlp(x, prefix + "info_proj_query.").mean(dim=1)
    pp = mlp(positive, prefix + "info_proj_positive.").mean(dim=1)
    return info_nce(pq, pp, temperature)


# --------------------------------------------------------------------------- K4
def ct_masks(mode, depth, output=None, w=0.2, coef=1.0):
    """Positive / negative boolean masks of CT_Regress (models/contrastive.py:17-31),
    CT_Single (:74-86) and CT_Multi (:115-141).  Returns (pos, neg, l_dist)."""
    n = depth.shape[0]
    eye = torch.eye(n, dtype=torch.bool, device=depth.device)
    if mode == "regress":
        l = depth.reshape(n, -1).mean(dim=1, keepdim=True)
        o = output.reshape(n, -1).mean(dim=1, keepdim=True)
        l_dist = (l - l.T).abs()
        p_dist = (o - o.T).abs()
        close = l_dist.le(w)
        return close & ~eye, (~close) & p_dist.le(w), l_dist
    if mode == "single":
        l = depth.reshape(n, 1)
        l_dist = (l - l.T).abs()
        same = l_dist.eq(0)
        return same & ~eye, ~same, l_dist
    if mode == "multi":
        d = depth.reshape(n, -1)
        c = d.shape[1]
        agree = (d[:, None, :] == d[None, :, :]).sum(-1).to(torch.float32) / c
        pos = agree.ge(coef / c)
        return pos & ~eye, ~pos, agree
    raise ValueError(mode)


def _ct_core(feature, pos, neg, pushing_w, denom, t):
    """Shared tail of the three CT losses (models/contrastive.py:33-35,45-55 / 88-110 /
    143-167).  Note the quirk: exp(pos).sum(1) runs over ALL columns and pos is 0 where
    the mask is 0, so every non-positive column adds exp(0)=1 to the partition sum."""
    f = F.normalize(feature.reshape(feature.shape[0], -1), dim=1)
    prod = (f @ f.T) / t
    pos_s = prod * pos
    neg_s = prod * neg
    neg_exp = (pushing_w * torch.exp(neg_s) * neg).sum(1)
    flag = neg.sum(1).bool()
    z = torch.exp(pos_s).sum(1) + neg_exp
    loss = ((-torch.log(torch.exp(pos_s) / z.unsqueeze(-1)) * pos).sum(1) / denom)
    return (loss * flag).unsqueeze(-1).mean()


def ct_regress(feature, depth, output, weights=None, w=0.2, t=0.07, e=0.01):
    """CT_Regress (ConR), models/contrastive.py:3-59."""
    pos, neg, l_dist = ct_masks("regress", depth, output, w)
    if weights is None:
        weights = torch.ones_like(l_dist)
    wrow = weights.reshape(weights.shape[0], -1).mean(dim=1, keepdim=True)
    pushing_w = l_dist * wrow * e
    denom = l_dist.le(w).sum(1)                 # counts the diagonal (:51)
    return _ct_core(feature, pos, neg, pushing_w, denom, t)


def ct_single(feature, depth, output=None, weights=None, t=0.07):
    """CT_Single (SupCon-style), models/contrastive.py:62-112.  ``weights`` broadcasts
    as passed (:94,97): scalar/(1,) -> all, (N,) -> per COLUMN, (N,1) -> per ROW."""
    pos, neg, _ = ct_masks("single", depth)
    if weights is None:
        weights = torch.tensor([1], device=feature.device)
    denom = pos.sum(1)
    denom = torch.where(denom == 0, torch.ones_like(denom), denom)
    return _ct_core(feature, pos, neg, weights.to(feature.device), denom, t)


def ct_multi(feature, depth, output=None, weights=None, t=0.07, coef=1.0):
    """CT_Multi, models/contrastive.py:114-169 (the O(N²) Python loop vectorised)."""
    pos, neg, _ = ct_masks("multi", depth, coef=coef)
    pw = torch.ones((), device=feature.device) if weights is None else weights.to(feature.device)
    denom = pos.sum(1)
    denom = torch.where(denom == 0, torch.ones_like(denom), denom)
    return _ct_core(feature, pos, neg, pw, denom, t)


# --------------------------------------------------------------------------- K5
def calibrate_mean_var(matrix, m1, v1, m2, v2, clip_min=0.1, clip_max=10):
    """utils/util.py:159-169 (three data-dependent branches; the middle one is in place)."""
    if torch.sum(v1) < 1e-10:
        return matrix
    if (v1 == 0.).any():
        valid = v1 != 0.
        factor = torch.clamp(v2[valid] / v1[valid], clip_min, clip_max)
        matrix[:, valid] = (matrix[:, valid] - m1[valid]) * torch.sqrt(factor) + m2[valid]
        return matrix
    factor = torch.clamp(v2 / v1, clip_min, clip_max)
    return (matrix - m1) * torch.sqrt(factor) + m2


def fds_label_bins(labels, min_value, bin_width):
    """models/fds.py:120-125,160-164: per-sample ``int((value-min)//bin_width)`` where
    ``value`` is a 0-dim fp32 tensor and min/bin_width are host floats, i.e. an fp32
    subtract followed by torch's fp32 floor_divide.  Vectorised; returns int64 (N,)."""
    l0 = labels[:, 0] if labels.dim() > 1 else labels
    l0 = l0.detach().to("cpu", torch.float32)
    return torch.floor_divide(l0 - float(min_value), float(bin_width)).to(torch.int64)


def _fds_groups(bins, bucket_start, bucket_num):
    """The (bucket, row-selector) pairs the loops at models/fds.py:133-141 / :166-189
    visit: only bins PRESENT in the batch; edge buckets absorb the tails (Q10)."""
    out = []
    for label in torch.unique(bins).tolist():
        if label > bucket_num - 1 or label < bucket_start:
            continue
        if label == bucket_start:
            sel = bins <= label
        elif label == bucket_num - 1:
            sel = bins >= label
        else:
            sel = bins == label
        out.append((int(label - bucket_start), sel))
    return out


def fds_smooth(features, labels, epoch, st, cfg):
    """FDS.smooth, models/fds.py:157-190.  ``st`` = dict of the FDS buffers,
    ``cfg`` = dict(min_value, bin_width, bucket_num, bucket_start, start_smooth).
    Writes into ``features`` in place and returns it (Q11)."""
    if epoch < cfg["start_smooth"]:
        return features
    bins = fds_label_bins(labels, cfg["min_value"], cfg["bin_width"]).to(features.device)
    for b, sel in _fds_groups(bins, cfg["bucket_start"], cfg["bucket_num"]):
        features[sel] = calibrate_mean_var(features[sel],
                                           st["running_mean_last_epoch"][b],
                                           st["running_var_last_epoch"][b],
                                           st["smoothed_mean_last_epoch"][b],
                                           st["smoothed_var_last_epoch"][b])
    return features


def fds_update_running_stats(features, labels, epoch, st, cfg):
    """FDS.update_running_stats, models/fds.py:116-155."""
    if epoch < float(st["epoch"]):
        return
    bins = fds_label_bins(labels, cfg["min_value"], cfg["bin_width"]).to(features.device)
    momentum = cfg.get("momentum", 0.9)
    for b, sel in _fds_groups(bins, cfg["bucket_start"], cfg["bucket_num"]):
        cur = features[sel]
        n = cur.size(0)
        mean = cur.mean(0)
        var = cur.var(0, unbiased=(n != 1))
        st["num_samples_tracked"][b] += n
        factor = momentum if momentum is not None else 1 - n / float(st["num_samples_tracked"][b])
        if epoch == cfg["start_update"]:
            factor = 0
        st["running_mean"][b] = (1 - factor) * mean + factor * st["running_mean"][b]
        st["running_var"][b] = (1 - factor) * var + factor * st["running_var"][b]


def fds_kernel_window(kernel="gaussian", ks=5, sigma=1):
    """FDS._get_kernel_window, models/fds.py:69-84 (gaussian via an explicit
    restatement of scipy.ndimage.gaussian_filter1d on a unit impulse with the default
    'reflect' boundary and truncate=4)."""
    half = (ks - 1) // 2
    if kernel == "gaussian":
        r = int(4.0 * sigma + 0.5)
        xs = torch.arange(-r, r + 1, dtype=torch.float64)
        w = torch.exp(-0.5 * (xs / sigma) ** 2)
        w = w / w.sum()
        base = torch.zeros(ks, dtype=torch.float64)
        base[half] = 1.0
        # scipy 'reflect' == symmetric (d c b a | a b c d | d c b a)
        n = ks
        idx = torch.arange(-r, n + r)
        period = 2 * n
        m = idx % period
        m = torch.where(m >= n, period - 1 - m, m)
        padded = base.to(torch.float32).to(torch.float64)[m]
        out = torch.stack([(padded[i:i + 2 * r + 1] * w).sum() for i in range(n)])
        win = out / out.sum()
    elif kernel == "triang":
        xs = torch.arange(1, ks + 1, dtype=torch.float64)
        tri = 1 - (xs - (ks + 1) / 2).abs() / ((ks + 1) / 2) if ks % 2 == 1 else \
            1 - (xs - (ks + 1) / 2).abs() / (ks / 2)
        win = tri / tri.sum()
    elif kernel == "laplace":
        xs = torch.arange(-half, half + 1, dtype=torch.float64)
        lap = torch.exp(-xs.abs() / sigma) / (2.0 * sigma)
        win = lap / lap.sum()
    else:
        raise AssertionError(kernel)
    return win.to(torch.float32)


def fds_update_last_epoch_stats(epoch, st, window):
    """FDS.update_last_epoch_stats/_update_last_epoch_stats, models/fds.py:86-99,110-114.
    The reference ALIASES running_* as running_*_last_epoch (Q9); the dict does the same."""
    if epoch != float(st["epoch"]) + 1:
        return
    st["epoch"] += 1
    st["running_mean_last_epoch"] = st["running_mean"]
    st["running_var_last_epoch"] = st["running_var"]
    half = (window.numel() - 1) // 2

    def smooth(t):       # (nb, D): reflect-pad + conv along the bucket axis
        x = t.t().unsqueeze(1)                                   # (D,1,nb)
        x = F.pad(x, (half, half), mode="reflect")
        y = F.conv1d(x, window.view(1, 1, -1).to(t))
        return y.squeeze(1).t().contiguous()

    st["smoothed_mean_last_epoch"] = smooth(st["running_mean_last_epoch"])
    st["smoothed_var_last_epoch"] = smooth(st["running_var_last_epoch"])


# ------------------------------------------------------------------ task losses
def mse_loss(pred, target):
    """nn.MSELoss registered for regression at models/nnmodel.py:27."""
    return F.mse_loss(pred, target)


# --------------------------------------------------------------------------- f2 cross-modal fusion
def _bert_ln(x, w, b, eps):
    """BertLayerNorm, TF style, epsilon inside the square root (models/mm_module.py:320-333)."""
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def cross_layer(s1, s2, mask2, p, prefix, heads=16, eps=1e-12, keeps=None, attn_dropout=0.0, dropout=0.0):
    """BertCrossAttentionLayer (models/mm_module.py:607-620): BertCoAttention (:493-522: q from s1, k / v from s2, scores /
    sqrt(d) + additive mask, softmax, dropout) -> BertSelfOutput (:525-536: dense, dropout, LayerNorm(x + s1)) ->
    BertIntermediate (:563-575: dense + erf GELU) -> BertOutput (:578-589).  mask2 (B, L2) 1 = attend; the additive form
    (1 - mask) * -10000 is models/mm_model.py:393-394.  keeps = (attention, attention-output, output) keep masks or None."""
    B, L1, D = s1.shape
    L2 = s2.shape[1]
    hd = D // heads
    lin = lambda x, n: F.linear(x, p[prefix + n + ".weight"], p[prefix + n + ".bias"])
    split = lambda t, L: t.view(B, L, heads, hd).permute(0, 2, 1, 3)
    q, k, v = split(lin(s1, "attention.self.query"), L1), split(lin(s2, "attention.self.key"), L2), split(lin(s2, "attention.self.value"), L2)
    sc = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(hd)
    sc = sc + ((1.0 - mask2.to(sc.dtype)) * -10000.0)[:, None, None, :]
    pr = torch.softmax(sc, dim=-1)
    pr = _drop(pr, attn_dropout, None if keeps is None else keeps[0])
    ctx = torch.matmul(pr, v).permute(0, 2, 1, 3).contiguous().view(B, L1, D)
    a = _drop(lin(ctx, "attention.output.dense"), dropout, None if keeps is None else keeps[1])
    a = _bert_ln(a + s1, p[prefix + "attention.output.LayerNorm.weight"], p[prefix + "attention.output.LayerNorm.bias"], eps)
    z = lin(a, "intermediate.dense")
    u = z * 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))                      # models/mm_module.py:204-211
    o = _drop(lin(u, "output.dense"), dropout, None if keeps is None else keeps[2])
    return _bert_ln(o + a, p[prefix + "output.LayerNorm.weight"], p[prefix + "output.LayerNorm.bias"], eps)


def cross_modal(text_embeddings, graph_embeddings, text_mask, graph_mask, p, heads=16, eps=1e-12, prefix=""):
    """CrossAttentionModel.forward (models/mm_model.py:386-406) with one layer per direction and dropout off:
    returns (text_to_graph, graph_to_text)."""
    g2t = cross_layer(graph_embeddings, text_embeddings, text_mask, p, prefix + "graph_attention.layer.0.", heads, eps)
    t2g = cross_layer(text_embeddings, graph_embeddings, graph_mask, p, prefix + "text_attention.layer.0.", heads, eps)
    return t2g, g2t


def fuse_pool(cross_txt, cross, img_mask, attention_mask):
    """models/mm_model.py:572-576: zero the rows outside the masks, concatenate, sum over tokens, divide by the valid counts."""
    a = cross_txt * img_mask[..., None].to(cross_txt.dtype)
    b = cross * attention_mask[..., None].to(cross.dtype)
    final = torch.cat((a, b), dim=1)
    return final.sum(dim=1) / (img_mask.sum(dim=1).view(-1, 1) + attention_mask.sum(dim=1).view(-1, 1))


# --------------------------------------------------------------------------- f4 ChemBERTa (HF RoBERTa encoder)
def roberta_encoder(input_ids, attention_mask, p, heads, n_layers, eps=1e-12, pad_idx=1, prefix=""):
    """``self.bert(input_ids, attention_mask, return_dict=True)[0]`` (models/mm_model.py:475,562): Hugging Face RobertaModel
    (third-party ``transformers``, unpinned by the reference), dropout off.  Embeddings = word + token-type 0 + position
    (position id = pad_idx + running count of non-padding tokens) -> LayerNorm; every layer is the post-LN block of
    ``cross_layer`` with s1 = s2 (self-attention; HF masks padded keys with the dtype minimum instead of -10000: both give
    exactly zero probability in fp32)."""
    keep = input_ids.ne(pad_idx).int()
    pos = (torch.cumsum(keep, dim=1) * keep).long() + pad_idx
    e = prefix + "embeddings."
    x = p[e + "word_embeddings.weight"][input_ids] + p[e + "token_type_embeddings.weight"][torch.zeros_like(input_ids)] \
        + p[e + "position_embeddings.weight"][pos]
    x = _bert_ln(x, p[e + "LayerNorm.weight"], p[e + "LayerNorm.bias"], eps)
    for i in range(n_layers):
        x = cross_layer(x, x, attention_mask, p, prefix + "encoder.layer.%d." % i, heads, eps)
    return x

