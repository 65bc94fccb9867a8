"""The slice of Uni-Core's API the reference imports (models/transformers.py:11,
models/mm_model.py:13-16), re-implemented on the mmdti_b200 kernels so the drop-in
does not depend on Uni-Core: LayerNorm, init_bert_params, get_activation_fn, Dictionary,
SelfMultiheadAttention, TransformerEncoderLayer (same constructor arguments, parameter
names and return conventions)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import config, ops
from .._lib import MMDTIError

UNIMOL_DICT = ["[PAD]", "[CLS]", "[SEP]", "[UNK]", "C", "N", "O", "S", "H", "Cl", "F", "Br", "I", "Si",
               "P", "B", "Na", "K", "Al", "Ca", "Sn", "As", "Hg", "Fe", "Zn", "Cr", "Se", "Gd", "Au", "Li"]


def get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    if activation == "tanh":
        return torch.tanh
    if activation == "linear":
        return lambda x: x
    raise RuntimeError("--activation-fn {} not supported".format(activation))


class LayerNorm(nn.LayerNorm):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)

    def forward(self, x):
        D = self.normalized_shape[-1]
        if x.is_cuda and self.elementwise_affine and len(self.normalized_shape) == 1 and D % 4 == 0 and D <= 1024:
            return ops.LayerNormFn.apply(x, self.weight, self.bias, self.eps)
        return F.layer_norm(x.float(), self.normalized_shape, self.weight, self.bias, self.eps)


def init_bert_params(module):
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()


class Dictionary:
    """Symbol table in file order (one symbol per line)."""

    def __init__(self, *, bos="[CLS]", pad="[PAD]", eos="[SEP]", unk="[UNK]"):
        self.bos_word, self.pad_word, self.eos_word, self.unk_word = bos, pad, eos, unk
        self.symbols, self.indices = [], {}

    def __len__(self):
        return len(self.symbols)

    def index(self, sym):
        return self.indices.get(sym, self.indices.get(self.unk_word))

    def add_symbol(self, word, is_special=False):
        if word not in self.indices:
            self.indices[word] = len(self.symbols)
            self.symbols.append(word)
        return self.indices[word]

    def bos(self):
        return self.index(self.bos_word)

    def pad(self):
        return self.index(self.pad_word)

    def eos(self):
        return self.index(self.eos_word)

    def unk(self):
        return self.index(self.unk_word)

    @classmethod
    def load(cls, path):
        d = cls()
        with open(path, "r", encoding="utf-8") as fh:
            for line in fh:
                tok = line.strip().split()
                if tok:
                    d.add_symbol(tok[0])
        return d

    @classmethod
    def unimol_default(cls):
        d = cls()
        for s in UNIMOL_DICT:
            d.add_symbol(s)
        return d


def _lin(x, layer, dt):
    return F.linear(x.to(dt), layer.weight.to(dt), None if layer.bias is None else layer.bias.to(dt))


class SelfMultiheadAttention(nn.Module):
    """in_proj / out_proj around the pair-biased attention kernel (head_dim must be 8)."""

    def __init__(self, embed_dim, num_heads, dropout=0.1, bias=True, scaling_factor=1):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = (self.head_dim * scaling_factor) ** -0.5
        self.in_proj = nn.Linear(embed_dim, embed_dim * 3, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

    def forward(self, query, key_padding_mask=None, attn_bias=None, return_attn=False, inplace_pair=False):
        if self.head_dim != 8:
            raise MMDTIError("mmdti_b200 pair attention is built for head_dim 8 (Uni-Mol: 512/64); got %d" % self.head_dim)
        B, L, D = query.shape
        H = self.num_heads
        dt = config.act_dtype()
        pdt = config.pair_dtype()
        if attn_bias is None:
            attn_bias = torch.zeros((B * H, L, L), device=query.device, dtype=pdt)
        # internal callers hand over the padded (B,H,L,Lp) layout; the reference's dense
        # (B*H,L,L) bias is padded here and the returned scores are un-padded again
        padded_in = attn_bias.dim() == 4 and attn_bias.shape[-1] == ops.pair_ld(L) and attn_bias.dtype == pdt
        if padded_in:
            pair = attn_bias.contiguous()
        else:
            pair = ops.PairPadFn.apply(attn_bias.reshape(B * H, L, L), B, H, L, pdt)
        if key_padding_mask is not None:
            pair = pair.clone()
            pair = ops.pair_mask_fill_(pair, key_padding_mask)
        qkv = _lin(query, self.in_proj, dt).reshape(B * L, 3 * D)
        p = self.dropout if self.training else 0.0
        o, scores = ops.pair_attention(qkv, pair, B, H, L, self.scaling, p, ops.next_seed() if p > 0 else 0,
                                       inplace_pair)
        o = _lin(o, self.out_proj, dt).view(B, L, D)
        if not return_attn:
            return o
        if not padded_in:
            scores = scores[..., :L].reshape(B * H, L, L)
        # Uni-Core also returns the (B*H,L,L) probabilities; they are never materialised here
        return o, scores, None


class TransformerEncoderLayer(nn.Module):
    """Pre-LN (or post-LN) encoder layer; with return_attn=True also returns the pre-softmax
    scores (QK^T*scale + bias), which TransformerEncoderWithPair feeds to the next layer."""

    def __init__(self, embed_dim=768, ffn_embed_dim=3072, attention_heads=8, dropout=0.1, attention_dropout=0.1,
                 activation_dropout=0.0, activation_fn="gelu", post_ln=False):
        super().__init__()
        self.embed_dim, self.attention_heads = embed_dim, attention_heads
        self.attention_dropout = attention_dropout
        self.dropout, self.activation_dropout = dropout, activation_dropout
        self.activation_fn = get_activation_fn(activation_fn)
        self.self_attn = SelfMultiheadAttention(embed_dim, attention_heads, dropout=attention_dropout)
        self.self_attn_layer_norm = LayerNorm(embed_dim)
        self.fc1 = nn.Linear(embed_dim, ffn_embed_dim)
        self.fc2 = nn.Linear(ffn_embed_dim, embed_dim)
        self.final_layer_norm = LayerNorm(embed_dim)
        self.post_ln = post_ln

    def _fusable(self, x, attn_bias, padding_mask, return_attn):
        return (return_attn and not self.post_ln and self.activation_fn is F.gelu and padding_mask is None
                and (self.activation_dropout == 0.0 or not self.training) and self.self_attn.head_dim == 8
                and attn_bias is not None and attn_bias.dim() == 4 and x.is_cuda
                and attn_bias.shape[-1] == ops.pair_ld(x.shape[1]) and attn_bias.dtype == config.pair_dtype()
                and x.shape[-1] % 4 == 0 and x.shape[-1] <= 1024)

    def lowp_weights(self, dt):
        """low-precision copies of the GEMM operands (refreshed by the owning encoder once per step)."""
        sa = self.self_attn
        return [t.detach().to(dt) for t in (sa.in_proj.weight, sa.in_proj.bias, sa.out_proj.weight, sa.out_proj.bias,
                                            self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)]

    def forward(self, x, attn_bias=None, padding_mask=None, return_attn=False, inplace_pair=False, lowp=None,
                chain_in=None, next_ln=None, link_in=None, link_out=None):
        """chain_in / next_ln (fused path only, used by TransformerEncoderWithPair): ``chain_in`` = (h1, stats) of this
        layer's LayerNorm-1 already computed by the previous layer, ``next_ln`` = the next layer's LayerNorm-1 module;
        with next_ln the result is (x, scores, None, (h_next, stats_next))."""
        dt = config.act_dtype()
        if self._fusable(x, attn_bias, padding_mask, return_attn):
            # single autograd node with a hand-written backward (ops.EncoderLayerFn)
            B, L, D = x.shape
            sa = self.self_attn
            train = self.training
            p_attn = self.attention_dropout if train else 0.0
            p_drop = self.dropout if train else 0.0
            seeds = (ops.next_seed(), ops.next_seed(), ops.next_seed()) if train else (0, 0, 0)
            cfg = (B, self.attention_heads, L, sa.scaling, p_attn, p_drop, seeds, dt)
            if dt == torch.float32:
                lowp = None
            h1_in, st1_in = chain_in if chain_in is not None else (None, None)
            out = ops.EncoderLayerFn.apply(
                x, attn_bias.contiguous(), self.self_attn_layer_norm.weight, self.self_attn_layer_norm.bias,
                sa.in_proj.weight, sa.in_proj.bias, sa.out_proj.weight, sa.out_proj.bias,
                self.final_layer_norm.weight, self.final_layer_norm.bias, self.fc1.weight, self.fc1.bias,
                self.fc2.weight, self.fc2.bias, lowp, cfg, h1_in, st1_in,
                None if next_ln is None else next_ln.weight, None if next_ln is None else next_ln.bias, link_in, link_out)
            if next_ln is not None:
                return out[0], out[1], None, (out[2], out[3])
            return out[0], out[1], None
        residual = x
        if not self.post_ln:
            x = self.self_attn_layer_norm(x)
        x = self.self_attn(query=x, key_padding_mask=padding_mask, attn_bias=attn_bias, return_attn=return_attn,
                           inplace_pair=inplace_pair)
        if return_attn:
            x, attn_weights, attn_probs = x
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if self.post_ln:
            x = self.self_attn_layer_norm(x)
        residual = x
        if not self.post_ln:
            x = self.final_layer_norm(x)
        x = _lin(x, self.fc1, dt)
        x = self.activation_fn(x)
        x = F.dropout(x, p=self.activation_dropout, training=self.training)
        x = _lin(x, self.fc2, dt)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if self.post_ln:
            x = self.final_layer_norm(x)
        if not return_attn:
            return x
        return x, attn_weights, attn_probs
