"""Shared helpers for the tests."""
import torch


def slice_shapes(H, D, Fd, nl, K=128, n_tok=31):
    """state_dict shapes of embed_tokens + gbf + gbf_proj + encoder."""
    s = {"embed_tokens.weight": (n_tok, D),
         "gbf.means.weight": (1, K), "gbf.stds.weight": (1, K),
         "gbf.mul.weight": (n_tok * n_tok, 1), "gbf.bias.weight": (n_tok * n_tok, 1),
         "gbf_proj.linear1.weight": (K, K), "gbf_proj.linear1.bias": (K,),
         "gbf_proj.linear2.weight": (H, K), "gbf_proj.linear2.bias": (H,),
         "encoder.emb_layer_norm.weight": (D,), "encoder.emb_layer_norm.bias": (D,),
         "encoder.final_layer_norm.weight": (D,), "encoder.final_layer_norm.bias": (D,)}
    for i in range(nl):
        p = "encoder.layers.%d." % i
        s.update({p + "self_attn.in_proj.weight": (3 * D, D), p + "self_attn.in_proj.bias": (3 * D,),
                  p + "self_attn.out_proj.weight": (D, D), p + "self_attn.out_proj.bias": (D,),
                  p + "self_attn_layer_norm.weight": (D,), p + "self_attn_layer_norm.bias": (D,),
                  p + "fc1.weight": (Fd, D), p + "fc1.bias": (Fd,), p + "fc2.weight": (D, Fd), p + "fc2.bias": (D,),
                  p + "final_layer_norm.weight": (D,), p + "final_layer_norm.bias": (D,)})
    return s


def seeded(shape, seed, scale=1.0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device)


def cross_layer_shapes(D, Fd, prefix=""):
    """state_dict shapes of one BertCrossAttentionLayer (models/mm_module.py:607-620)."""
    s = {}
    for n in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
        s[prefix + n + ".weight"], s[prefix + n + ".bias"] = (D, D), (D,)
    s[prefix + "intermediate.dense.weight"], s[prefix + "intermediate.dense.bias"] = (Fd, D), (Fd,)
    s[prefix + "output.dense.weight"], s[prefix + "output.dense.bias"] = (D, Fd), (D,)
    for n in ("attention.output.LayerNorm", "output.LayerNorm"):
        s[prefix + n + ".weight"], s[prefix + n + ".bias"] = (D,), (D,)
    return s
