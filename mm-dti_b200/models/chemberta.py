"""The SMILES encoder of the second modality: ChemBERTa = a Hugging Face RoBERTa encoder, loaded by
the reference with ``AutoModel.from_pretrained(chemberta_dir)`` (models/mm_model.py:475) and called as
``self.bert(input_ids, attention_mask, return_dict=True)[0]`` (models/mm_model.py:562; models/encoder.py:548-572).

``RobertaModel`` below keeps the HF ``state_dict`` names (``embeddings.word_embeddings.weight``,
``encoder.layer.N.attention.self.query.weight`` ... ``pooler.dense.weight``), so HF checkpoints load unchanged, and runs every
layer as ONE autograd node on the mmdti kernels: a RoBERTa layer is the post-LN block of the cross-modal fusion with s1 = s2
(ops_cross.CrossLayerFn: tcgen05 GEMMs with fused epilogues + csrc/cross_attn.cu).  bf16 mode needs hidden % 64 == 0,
hidden <= 512 and head_dim 32 or 64 (the reference's ChemBERTa output feeds 512-d heads, models/mm_model.py:493);
``supported(config)`` says whether a configuration qualifies -- ``models.encoder.ChembertaEncoder`` falls back to the stock HF
module otherwise, loudly."""
import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops, ops_cross
from .cross_modal import BertCrossAttentionLayer, _Norm


def supported(cfg):
    hd = cfg.hidden_size // cfg.num_attention_heads
    return (cfg.hidden_size % 64 == 0 and cfg.hidden_size <= 512 and cfg.intermediate_size % 64 == 0 and hd in (32, 64)
            and cfg.hidden_act == "gelu" and getattr(cfg, "position_embedding_type", "absolute") == "absolute")


class RobertaEmbeddings(nn.Module):
    """word + token-type(0) + position embeddings, LayerNorm, dropout.  Position ids follow RoBERTa: padding_idx + the running
    count of non-padding tokens (padding positions stay at padding_idx)."""

    def __init__(self, cfg):
        super().__init__()
        self.padding_idx = cfg.pad_token_id
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = _Norm(cfg.hidden_size, cfg.layer_norm_eps)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size, padding_idx=cfg.pad_token_id)
        self.dropout = nn.Dropout(cfg.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None):
        keep = input_ids.ne(self.padding_idx).int()
        position_ids = (torch.cumsum(keep, dim=1) * keep).long() + self.padding_idx
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        x = self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids) + self.position_embeddings(position_ids)
        x = ops.LayerNormFn.apply(x, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.variance_epsilon)
        return ops_cross.flat_dropout(x, self.dropout.p, self.training)


class RobertaEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([BertCrossAttentionLayer(cfg) for _ in range(cfg.num_hidden_layers)])

    def forward(self, x, attention_mask):
        for i, layer in enumerate(self.layer):          # self-attention: queries, keys and values from the same stream;
            x = layer(x, x, attention_mask, out_f32=i == len(self.layer) - 1)      # fp32 only out of the last layer
        return x


class RobertaPooler(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dense = nn.Linear(cfg.hidden_size, cfg.hidden_size)

    def forward(self, x):
        return torch.tanh(self.dense(x[:, 0]))


class RobertaOutput(tuple):
    """(last_hidden_state, pooler_output) with the attribute names of HF's BaseModelOutputWithPooling."""

    @property
    def last_hidden_state(self):
        return self[0]

    @property
    def pooler_output(self):
        return self[1]


class RobertaModel(nn.Module):
    def __init__(self, cfg, add_pooling_layer=True):
        super().__init__()
        if not supported(cfg):
            raise ValueError("RobertaModel on the mmdti kernels needs hidden % 64 == 0, hidden <= 512, head_dim 32 or 64, gelu, absolute "
                             "positions (got hidden %d, heads %d, act %s)" % (cfg.hidden_size, cfg.num_attention_heads, cfg.hidden_act))
        self.config = cfg
        self.embeddings = RobertaEmbeddings(cfg)
        self.encoder = RobertaEncoder(cfg)
        self.pooler = RobertaPooler(cfg) if add_pooling_layer else None
        self.apply(self._init)

    def _init(self, m):
        std = getattr(self.config, "initializer_range", 0.02)
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, std)
            m.bias.data.zero_()
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(0.0, std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()

    def forward(self, input_ids, attention_mask=None, token_type_ids=None, return_dict=True):
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        x = self.embeddings(input_ids, token_type_ids)
        x = self.encoder(x, attention_mask)
        pooled = self.pooler(x) if self.pooler is not None else None
        return RobertaOutput((x, pooled))

    @classmethod
    def from_config(cls, cfg):
        return cls(cfg)

    @classmethod
    def from_pretrained(cls, path):
        """A local HF checkpoint directory: config.json + model.safetensors | pytorch_model.bin (a ``roberta.`` prefix and
        language-model heads, as saved by RobertaForMaskedLM, are tolerated)."""
        from transformers import RobertaConfig
        with open(os.path.join(path, "config.json")) as fh:
            cfg = RobertaConfig(**{k: v for k, v in json.load(fh).items() if k not in ("architectures", "model_type", "transformers_version")})
        model = cls(cfg)
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
        sd = {(k[len("roberta."):] if k.startswith("roberta.") else k): v for k, v in sd.items()}
        own = model.state_dict()
        missing = [k for k in own if k not in sd and not k.startswith("pooler.")]
        if missing:
            raise KeyError("checkpoint %s lacks %s" % (path, missing[:4]))
        model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
        return model
