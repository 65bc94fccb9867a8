"""Drop-in for the reference's models/contrastive.py: CT_Regress (ConR), CT_Single
(SupCon-style) and CT_Multi (multi-label) with the reference's signatures, on the fused
similarity kernels (ops_sim.ContrastiveFn): normalise + Gram tiles on the tensor cores +
label/distance masks and pushing weights evaluated on the fly + weighted exp-sums, one pass
forward, one recompute pass backward.  No N x N temporaries, no Python loop over the diagonal
(contrastive.py:30-31) and no O(N^2) host loop for the multi-label overlap (:115-123).

Every function takes an extra keyword ``dp`` (dist.DataParallelCtx): the local batch then is
this rank's slice of the global batch and negatives come from all ranks."""
import torch

from .. import ops_sim


def CT_Regress(feature, depth, output, weights=None, w=0.2, t=0.07, e=0.01, dp=None):
    """ConR: positives |y_i - y_j| <= w (diagonal removed), negatives |y_i - y_j| > w and
    |yhat_i - yhat_j| <= w, pushing weight |y_i - y_j| * mean(weights_i) * e (contrastive.py:3-59)."""
    return ops_sim.ct_regress(feature, depth, output, weights=weights, w=w, t=t, e=e, dp=dp)


def CT_Single(feature, depth, output, weights=torch.tensor([1]), w=0.2, t=0.07, e=0.2, lamda=1, dp=None):
    """Positives = same label (diagonal removed), negatives = different label; ``w, e, lamda,
    output`` are accepted and ignored like in the reference (contrastive.py:62-112)."""
    return ops_sim.ct_single(feature, depth, output, weights=weights, t=t, dp=dp)


def CT_Multi(feature, depth, output, weights=None, w=0.2, t=0.07, e=0.2, coef=1, dp=None):
    """Positives = fraction of equal label columns >= coef / n_classes (contrastive.py:114-169)."""
    return ops_sim.ct_multi(feature, depth, output, weights=weights, t=t, coef=coef, dp=dp)
