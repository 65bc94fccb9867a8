"""TransformerEncoderWithPair — drop-in for the reference's models/transformers.py:14-183
(same constructor, forward signature, return tuple and state_dict names), running on the
mmdti_b200 kernels.  The pair tensor is carried as (B,H,L,L) in config.pair_dtype()."""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import config, ops
from .unicore_compat import LayerNorm, TransformerEncoderLayer


class TransformerEncoderWithPair(nn.Module):
    def __init__(
        self,
        encoder_layers: int = 6,
        embed_dim: int = 768,
        ffn_embed_dim: int = 3072,
        attention_heads: int = 8,
        emb_dropout: float = 0.1,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.0,
        max_seq_len: int = 256,
        activation_fn: str = "gelu",
        post_ln: bool = False,
        no_final_head_layer_norm: bool = False,
    ) -> None:
        super().__init__()
        self.emb_dropout = emb_dropout
        self.max_seq_len = max_seq_len
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        self.emb_layer_norm = LayerNorm(self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim) if not post_ln else None
        self.final_head_layer_norm = LayerNorm(attention_heads) if not no_final_head_layer_norm else None
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(embed_dim=self.embed_dim, ffn_embed_dim=ffn_embed_dim,
                                    attention_heads=attention_heads, dropout=dropout,
                                    attention_dropout=attention_dropout, activation_dropout=activation_dropout,
                                    activation_fn=activation_fn, post_ln=post_ln)
            for _ in range(encoder_layers)
        ])
        # MM-DTI discards outputs 1..4 (models/mm_model.py:559); set False to skip producing them
        self.pair_outputs = True
        self.chain_layers = True      # fuse dropout+residual of layer i with LayerNorm-1 of layer i+1

    def forward(self, emb: torch.Tensor, attn_mask: Optional[torch.Tensor] = None,
                padding_mask: Optional[torch.Tensor] = None):
        """emb (B,L,D); attn_mask (B*H,L,L) pair bias — MUTATED IN PLACE with -inf at padded key
        columns exactly like models/transformers.py:122-132; padding_mask (B,L) bool or None.
        Returns (x, pair (B,L,L,H), delta_pair (B,L,L,H), x_norm, delta_pair_norm)."""
        bsz, seq_len = emb.size(0), emb.size(1)
        H = self.attention_heads
        assert attn_mask is not None
        if padding_mask is not None:
            if not attn_mask.is_contiguous():
                raise ValueError("attn_mask must be contiguous (B*H, L, L)")
            attn_mask = ops.pair_mask_fill_(attn_mask, padding_mask)      # the caller's tensor, in place (Q1)
        pair = ops.PairPadFn.apply(attn_mask.reshape(bsz * H, seq_len, seq_len), bsz, H, seq_len, config.pair_dtype())
        return self.forward_padded(emb, pair, padding_mask)

    def _gemm_params(self):
        src = []
        for layer in self.layers:
            sa = layer.self_attn
            src += [sa.in_proj.weight, sa.in_proj.bias, sa.out_proj.weight, sa.out_proj.bias, layer.fc1.weight,
                    layer.fc1.bias, layer.fc2.weight, layer.fc2.bias]
        return src

    def use_external_lowp(self):
        """Hand the bf16 shadows of the GEMM operands to an optimizer that refreshes them while it updates the
        fp32 master weights (mmdti_b200.optim.FusedAdam(shadows=...)): forward then skips its per-step cast pass.
        Returns {parameter: bf16 shadow}.  Call again after loading new weights."""
        src = self._gemm_params()
        shadows = [p.detach().to(torch.bfloat16) for p in src]
        self._lowp_cache = shadows
        self._lowp_external = True
        return dict(zip(src, shadows))

    def _lowp_weights(self):
        """bf16 copies of every layer's GEMM operands, refreshed with one multi-tensor copy per step."""
        dt = config.act_dtype()
        if dt == torch.float32:
            return None
        if getattr(self, "_lowp_external", False) and dt == torch.bfloat16:
            return self._lowp_cache
        src = []
        for layer in self.layers:
            sa = layer.self_attn
            src += [sa.in_proj.weight, sa.in_proj.bias, sa.out_proj.weight, sa.out_proj.bias, layer.fc1.weight,
                    layer.fc1.bias, layer.fc2.weight, layer.fc2.bias]
        src = [t.detach() for t in src]
        cache = getattr(self, "_lowp_cache", None)
        if (cache is None or len(cache) != len(src) or cache[0].dtype != dt or cache[0].device != src[0].device
                or any(c.shape != s_.shape for c, s_ in zip(cache, src))):
            cache = [torch.empty_like(t, dtype=dt) for t in src]
            self._lowp_cache = cache
        torch._foreach_copy_(cache, src)
        return cache

    def forward_padded(self, emb, pair, padding_mask=None):
        """Same as forward() for a pair bias that already is in the library's padded (B,H,L,Lp)
        layout with the key-padding mask merged (-inf columns), e.g. straight from K1."""
        bsz, seq_len = emb.size(0), emb.size(1)
        H = self.attention_heads
        x = self.emb_layer_norm(emb)
        x = F.dropout(x, p=self.emb_dropout, training=self.training)
        if padding_mask is not None:
            x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))
        pair_first = pair
        lowp = self._lowp_weights()
        # cross-layer fusion: the final dropout+residual of layer i and LayerNorm-1 of layer i+1 run as one kernel
        # (forward and backward) whenever both layers take the fused path; `chain` carries (h1, statistics)
        chain = None
        link = None                  # dict shared by two adjacent layers of the chain (see ops.EncoderLayerFn)
        n_layers = len(self.layers)
        for i, layer in enumerate(self.layers):
            nxt = self.layers[i + 1] if i + 1 < n_layers else None
            fuse_next = (self.chain_layers and nxt is not None and layer._fusable(x, pair, None, True)
                         and isinstance(nxt.self_attn_layer_norm, torch.nn.LayerNorm))
            link_next = {} if fuse_next else None
            out = layer(x, padding_mask=None, attn_bias=pair, return_attn=True,
                        lowp=None if lowp is None else lowp[8 * i:8 * i + 8],
                        chain_in=chain, next_ln=nxt.self_attn_layer_norm if fuse_next else None,
                        link_in=link if chain is not None else None, link_out=link_next)
            x, pair = out[0], out[1]
            chain = out[3] if len(out) > 3 else None
            link = link_next

        if not self.pair_outputs:
            if self.final_layer_norm is not None:
                x = self.final_layer_norm(x)
            return x, None, None, None, None

        def norm_loss(t, eps=1e-10, tolerance=1.0):
            t = t.float()
            max_norm = t.shape[-1] ** 0.5
            norm = torch.sqrt(torch.sum(t ** 2, dim=-1) + eps)
            return F.relu((norm - max_norm).abs() - tolerance)

        def masked_mean(mask, value, dim=-1, eps=1e-10):
            return (torch.sum(mask * value, dim=dim) / (eps + torch.sum(mask, dim=dim))).mean()

        x_norm = norm_loss(x)
        if padding_mask is not None:
            token_mask = 1.0 - padding_mask.float()
        else:
            token_mask = torch.ones_like(x_norm)
        x_norm = masked_mean(token_mask, x_norm)
        if self.final_layer_norm is not None:
            x = self.final_layer_norm(x)
        pair_out, delta = ops.PairOutputsFn.apply(pair_first.contiguous(), pair.contiguous(), bsz, H, seq_len)
        pair_mask = token_mask[..., None] * token_mask[..., None, :]
        delta_norm = masked_mean(pair_mask, norm_loss(delta), dim=(-1, -2))
        if self.final_head_layer_norm is not None:
            delta = self.final_head_layer_norm(delta)
        return x, pair_out, delta, x_norm, delta_norm
