"""unicore.modules restated from the public Uni-Core algorithm:
LayerNorm, softmax_dropout, SelfMultiheadAttention, TransformerEncoderLayer,
init_bert_params.  Reference call sites: models/transformers.py:69-91,136-139;
models/mm_model.py:472.  UNPINNED third-party restatement."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .utils import get_activation_fn


class LayerNorm(nn.LayerNorm):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True):
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)


def softmax_dropout(x, dropout_prob, is_training=True, mask=None, bias=None, inplace=True):
    if mask is not None:
        x = x + mask
    if bias is not None:
        x = x + bias
    return F.dropout(F.softmax(x, dim=-1), p=dropout_prob, training=is_training)


def init_bert_params(module):
    if isinstance(module, nn.Linear):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        module.weight.data.normal_(mean=0.0, std=0.02)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()


class SelfMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.1, bias=True, scaling_factor=1):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = embed_dim, num_heads, dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim
        self.scaling = (self.head_dim * scaling_factor) ** -0.5
        self.in_proj = nn.Linear(embed_dim, embed_dim * 3, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

    def forward(self, query, key_padding_mask=None, attn_bias=None, return_attn=False):
        bsz, tgt_len, embed_dim = query.size()
        q, k, v = self.in_proj(query).chunk(3, dim=-1)

        def heads(t):
            return (t.view(bsz, -1, self.num_heads, self.head_dim).transpose(1, 2)
                    .contiguous().view(bsz * self.num_heads, -1, self.head_dim))

        q = heads(q) * self.scaling
        k, v = heads(k), heads(v)
        src_len = k.size(1)
        attn_weights = torch.bmm(q, k.transpose(1, 2))
        if key_padding_mask is not None:
            attn_weights = attn_weights.view(bsz, self.num_heads, tgt_len, src_len)
            attn_weights.masked_fill_(key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool),
                                      float("-inf"))
            attn_weights = attn_weights.view(bsz * self.num_heads, tgt_len, src_len)
        if not return_attn:
            attn = softmax_dropout(attn_weights, self.dropout, self.training, bias=attn_bias)
        else:
            attn_weights = attn_weights + attn_bias
            attn = softmax_dropout(attn_weights, self.dropout, self.training, inplace=False)
        o = torch.bmm(attn, v)
        o = (o.view(bsz, self.num_heads, tgt_len, self.head_dim).transpose(1, 2)
             .contiguous().view(bsz, tgt_len, embed_dim))
        o = self.out_proj(o)
        if not return_attn:
            return o
        return o, attn_weights, attn


class TransformerEncoderLayer(nn.Module):
    def __init__(self, embed_dim=768, ffn_embed_dim=3072, attention_heads=8, dropout=0.1,
                 attention_dropout=0.1, activation_dropout=0.0, activation_fn="gelu",
                 post_ln=False):
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

    def forward(self, x, attn_bias=None, padding_mask=None, return_attn=False):
        residual = x
        if not self.post_ln:
            x = self.self_attn_layer_norm(x)
        x = self.self_attn(query=x, key_padding_mask=padding_mask, attn_bias=attn_bias,
                           return_attn=return_attn)
        if return_attn:
            x, attn_weights, attn_probs = x
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if self.post_ln:
            x = self.self_attn_layer_norm(x)
        residual = x
        if not self.post_ln:
            x = self.final_layer_norm(x)
        x = self.fc1(x)
        x = self.activation_fn(x)
        x = F.dropout(x, p=self.activation_dropout, training=self.training)
        x = self.fc2(x)
        x = F.dropout(x, p=self.dropout, training=self.training)
        x = residual + x
        if self.post_ln:
            x = self.final_layer_norm(x)
        if not return_attn:
            return x
        return x, attn_weights, attn_probs
