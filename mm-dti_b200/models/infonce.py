"""Drop-in for the reference's models/infonce.py: ``InfoNCE`` module and ``info_nce`` function
(same constructor / call signatures, state_dict names ``info_proj_query.{0,2}``,
``info_proj_positive.{0,2}`` and the same ValueErrors, models/infonce.py:45-67).

The implicit-negatives loss MM_Model uses (models/mm_model.py:566-567 -> infonce.py:89-98) runs
on the fused similarity kernels (ops_sim.InfoNCEFn): normalise + similarity GEMM + row/column
log-sum-exp + diagonal in one pass, never materialising the N x N logits.  ``dp`` (a
dist.DataParallelCtx) makes every rank score its rows against the all-gathered global batch."""
import torch
import torch.nn.functional as F
from torch import nn

from .. import config, ops_gemm, ops_sim

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class InfoNCE(nn.Module):
    def __init__(self, bert_output_size, graph_ouput_size, temperature=0.1, reduction='mean', negative_mode='unpaired'):
        super().__init__()
        self.temperature = temperature
        self.reduction = reduction
        self.negative_mode = negative_mode
        self.orig_d_l = bert_output_size
        self.orig_d_av = graph_ouput_size
        self.d_l, self.d_av = 50, 50
        self.embed_dropout = 0.1
        self.training = True            # the reference sets it by hand (infonce.py:18); model.eval() overrides it
        self.info_proj_query = nn.Sequential(nn.Linear(self.orig_d_l, self.orig_d_l), nn.GELU(),
                                             nn.Linear(self.orig_d_l, self.d_l))
        self.info_proj_positive = nn.Sequential(nn.Linear(self.orig_d_av, self.orig_d_av), nn.GELU(),
                                                nn.Linear(self.orig_d_av, self.d_av))
        self.dp = None                  # set to a dist.DataParallelCtx for global-batch negatives

    @staticmethod
    def _mlp(seq, x):
        """512 -> 512 -> GELU -> 50 on every token of the batch (2 x 8 448 rows at config 2), then the UNMASKED mean over
        the sequence axis (infonce.py:24-33).  bf16 mode on a CUDA device: the first Linear + GELU (99 % of the flops) is
        the own tcgen05 GEMM with fused epilogue (ops_gemm.LinearGeluFn); the second Linear commutes with the mean
        (mean_l(W u_l + b) = W mean_l(u_l) + b), so it runs on the B pooled rows instead of on every token.
        fp32 validation mode / CPU: the stock composition, untouched."""
        lin0, act, lin2 = seq[0], seq[1], seq[2]
        B, L, D = x.shape
        if (x.is_cuda and not config.fp32_mode() and isinstance(act, nn.GELU) and D % 64 == 0 and lin0.out_features % 64 == 0
                and lin0.bias is not None):
            u = ops_gemm.linear_gelu(x.reshape(B * L, D), lin0.weight, lin0.bias)          # (B*L, 512) bf16
            pooled = u.view(B, L, -1).mean(dim=1, dtype=torch.float32)
            return F.linear(pooled, lin2.weight, lin2.bias)
        return torch.mean(seq(x), dim=1)

    def project(self, query, positive_key):
        """dropout(query) -> per-modality MLP -> UNMASKED mean over the sequence axis (infonce.py:24-33):
        padded positions do contribute, exactly like the reference."""
        q = F.dropout(query, p=self.embed_dropout, training=self.training)
        pq = torch.mean(q, dim=1) if self.orig_d_l == self.d_l else self._mlp(self.info_proj_query, q)
        pp = torch.mean(positive_key, dim=1) if self.orig_d_av == self.d_av else self._mlp(self.info_proj_positive, positive_key)
        return pq, pp

    def forward(self, query, positive_key, negative_keys=None):
        proj_query, proj_positive = self.project(query, positive_key)
        return info_nce(proj_query, proj_positive, negative_keys, temperature=self.temperature,
                        reduction=self.reduction, negative_mode=self.negative_mode, dp=self.dp)


def info_nce(query, positive_key, negative_keys=None, temperature=0.1, reduction='mean', negative_mode='unpaired', dp=None):
    if query.dim() != 2:
        raise ValueError('<query> must have 2 dimensions.')
    if positive_key.dim() != 2:
        raise ValueError('<positive_key> must have 2 dimensions.')
    if negative_keys is not None:
        if negative_mode == 'unpaired' and negative_keys.dim() != 2:
            raise ValueError("<negative_keys> must have 2 dimensions if <negative_mode> == 'unpaired'.")
        if negative_mode == 'paired' and negative_keys.dim() != 3:
            raise ValueError("<negative_keys> must have 3 dimensions if <negative_mode> == 'paired'.")
    if len(query) != len(positive_key):
        raise ValueError('<query> and <positive_key> must must have the same number of samples.')
    if negative_keys is not None and negative_mode == 'paired' and len(query) != len(negative_keys):
        raise ValueError("If negative_mode == 'paired', then <negative_keys> must have the same number of samples as <query>.")
    if query.shape[-1] != positive_key.shape[-1]:
        raise ValueError('Vectors of <query> and <positive_key> should have the same number of components.')
    if negative_keys is not None and query.shape[-1] != negative_keys.shape[-1]:
        raise ValueError('Vectors of <query> and <negative_keys> should have the same number of components.')

    if negative_keys is None and reduction in ("mean", "sum"):
        return ops_sim.info_nce_loss(query, positive_key, temperature, reduction, dp=dp)

    # Explicit-negatives modes and reduction='none' (infonce.py:71-88) are never reached by MM_Model
    # (SURVEY.md §3.4); they stay a stock composition of device ops, outside the fused path.
    query, positive_key, negative_keys = normalize(query, positive_key, negative_keys)
    if negative_keys is not None:
        positive_logit = torch.sum(query * positive_key, dim=1, keepdim=True)
        if negative_mode == 'unpaired':
            negative_logits = query @ transpose(negative_keys)
        else:
            negative_logits = (query.unsqueeze(1) @ transpose(negative_keys)).squeeze(1)
        logits = torch.cat([positive_logit, negative_logits], dim=1)
        labels = torch.zeros(len(logits), dtype=torch.long, device=query.device)
    else:
        logits = query @ transpose(positive_key)
        labels = torch.arange(len(query), device=query.device)
    return (F.cross_entropy(logits / temperature, labels, reduction=reduction)
            + F.cross_entropy(logits.T / temperature, labels, reduction=reduction)) / 2


def transpose(x):
    return x.transpose(-2, -1)


def normalize(*xs):
    return [None if x is None else F.normalize(x, dim=-1) for x in xs]
