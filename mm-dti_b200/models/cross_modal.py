"""Drop-in for the reference's cross-modal fusion block: ``CrossAttentionModel`` (models/mm_model.py:379-406) over two
``BertCrossEncoder`` stacks (models/mm_module.py:663-677) of post-LN ``BertCrossAttentionLayer`` (:607-620), plus the
masked mean pooling that follows it in ``MM_Model.forward`` (models/mm_model.py:571-576).

The module tree below exists for ONE reason: identical ``state_dict`` names (``text_attention.layer.0.attention.self.query.weight``
...), so reference checkpoints load with strict=True.  The sub-modules are parameter holders; the compute is one autograd
node per layer (ops_cross.CrossLayerFn: tcgen05 GEMMs with fused epilogues + csrc/cross_attn.cu)."""
import torch
import torch.nn as nn

from .. import config, ops, ops_cross


def crossmodal_config(**overrides):
    """The values of the reference's ``crossmodal_config()`` (models/mm_model.py:361-377) as a plain namespace."""
    import types
    cfg = types.SimpleNamespace(attention_probs_dropout_prob=0.2, gradient_checkpointing=False, hidden_act="gelu", hidden_dropout_prob=0.3,
                                hidden_size=512, initializer_range=0.02, intermediate_size=2048, layer_norm_eps=1e-12,
                                max_position_embeddings=512, num_attention_heads=16, num_hidden_layers=12,
                                position_embedding_type="absolute")
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


class _Holder(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):          # pragma: no cover
        raise RuntimeError("parameter holder: the enclosing layer runs the fused kernels")


class _Norm(_Holder):
    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.variance_epsilon = eps


class _Projections(_Holder):            # BertCoAttention: query / key / value (+ attention-probability dropout rate)
    def __init__(self, cfg):
        super().__init__()
        if cfg.hidden_size % cfg.num_attention_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (cfg.hidden_size, cfg.num_attention_heads))
        self.num_attention_heads = cfg.num_attention_heads
        self.query = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.key = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.value = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.dropout = nn.Dropout(cfg.attention_probs_dropout_prob)


class _DenseNorm(_Holder):              # BertSelfOutput / BertOutput: dense + dropout + residual LayerNorm
    def __init__(self, d_in, cfg):
        super().__init__()
        self.dense = nn.Linear(d_in, cfg.hidden_size)
        self.LayerNorm = _Norm(cfg.hidden_size, cfg.layer_norm_eps)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)


class _Attention(_Holder):              # BertCrossAttention
    def __init__(self, cfg):
        super().__init__()
        self.self = _Projections(cfg)
        self.output = _DenseNorm(cfg.hidden_size, cfg)


class _Intermediate(_Holder):           # BertIntermediate (exact-erf GELU, models/mm_module.py:204-211)
    def __init__(self, cfg):
        super().__init__()
        if cfg.hidden_act != "gelu":
            raise ValueError("cross-modal layer: only hidden_act='gelu' is built (the reference's crossmodal_config)")
        self.dense = nn.Linear(cfg.hidden_size, cfg.intermediate_size)


class BertCrossAttentionLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Intermediate(cfg)
        self.output = _DenseNorm(cfg.intermediate_size, cfg)

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask, out_f32=True):
        """``s2_attention_mask``: (B, L2) bool / 0-1 (1 = attend), or the reference's extended additive form
        (B, 1, 1, L2) with 0 / -10000 (models/mm_model.py:393-394).  ``out_f32`` (default): fp32 output like the reference
        under autocast (LayerNorm output), a fresh tensor because the caller zeroes masked rows IN PLACE (mm_model.py:572-573);
        False: the activation dtype, for consumers that read bf16 (the next layer of a stack, fuse_and_pool)."""
        m = s2_attention_mask
        if m.dim() == 4:
            m = m[:, 0, 0, :] > -5000.0
        att, out = self.attention, self.output
        train = self.training
        p_attn = att.self.dropout.p if train else 0.0
        p_hid = out.dropout.p if train else 0.0
        if att.output.dropout.p != out.dropout.p:
            raise ValueError("cross-modal layer: one hidden dropout rate per layer")
        seeds = (ops.next_seed(), ops.next_seed(), ops.next_seed()) if train else (0, 0, 0)
        cfg = (att.self.num_attention_heads, p_attn, p_hid, seeds, config.act_dtype(), att.output.LayerNorm.variance_epsilon, out_f32)
        return ops_cross.CrossLayerFn.apply(
            s1_hidden_states, s2_hidden_states, m, att.self.query.weight, att.self.query.bias, att.self.key.weight, att.self.key.bias,
            att.self.value.weight, att.self.value.bias, att.output.dense.weight, att.output.dense.bias, att.output.LayerNorm.weight,
            att.output.LayerNorm.bias, self.intermediate.dense.weight, self.intermediate.dense.bias, out.dense.weight, out.dense.bias,
            out.LayerNorm.weight, out.LayerNorm.bias, cfg)


class BertCrossEncoder(nn.Module):
    def __init__(self, config_, layer_num):
        super().__init__()
        self.layer = nn.ModuleList([BertCrossAttentionLayer(config_) for _ in range(layer_num)])
        for l in self.layer[1:]:                      # the reference deep-copies ONE initialised layer (mm_module.py:666-667)
            l.load_state_dict(self.layer[0].state_dict())

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask, output_all_encoded_layers=True, out_f32=True):
        outs = []
        for i, layer in enumerate(self.layer):
            last = i == len(self.layer) - 1
            s1_hidden_states = layer(s1_hidden_states, s2_hidden_states, s2_attention_mask,
                                     out_f32=out_f32 and (last or output_all_encoded_layers))
            if output_all_encoded_layers:
                outs.append(s1_hidden_states)
        if not output_all_encoded_layers:
            outs.append(s1_hidden_states)
        return outs


class CrossAttentionModel(nn.Module):
    """forward(text_embeddings, graph_embeddings, text_mask, graph_mask) -> (text_to_graph, graph_to_text), argument names as in
    the reference (which passes the graph tokens first, models/mm_model.py:571)."""

    def __init__(self, cross_cfg, num_layers=1):
        super().__init__()
        self.text_attention = BertCrossEncoder(cross_cfg, num_layers)
        self.graph_attention = BertCrossEncoder(cross_cfg, num_layers)
        self.dropout = nn.Dropout(cross_cfg.hidden_dropout_prob)

    def forward(self, text_embeddings, graph_embeddings, text_mask, graph_mask, out_f32=True):
        t = ops_cross.flat_dropout(text_embeddings, self.dropout.p, self.training)
        g = ops_cross.flat_dropout(graph_embeddings, self.dropout.p, self.training)
        graph_to_text = self.graph_attention(g, t, text_mask, out_f32=out_f32)[-1]
        text_to_graph = self.text_attention(t, g, graph_mask, out_f32=out_f32)[-1]
        return text_to_graph, graph_to_text

    def forward_pooled(self, text_embeddings, graph_embeddings, text_mask, graph_mask):
        """forward + the masked mean pooling of models/mm_model.py:572-576 without the fp32 copies of the two outputs."""
        t2g, g2t = self.forward(text_embeddings, graph_embeddings, text_mask, graph_mask, out_f32=False)
        return fuse_and_pool(t2g, g2t, text_mask, graph_mask)


def fuse_and_pool(cross_txt_output_layer, cross_output_layer, img_mask, attention_mask):
    """models/mm_model.py:572-576 in one kernel each way: masked rows zeroed, concatenation, sum over tokens, division by the
    number of valid tokens of both modalities -> (B, D) f32."""
    return ops_cross.masked_mean_pool(cross_txt_output_layer, img_mask, cross_output_layer, attention_mask)
