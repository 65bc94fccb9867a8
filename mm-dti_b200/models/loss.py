"""Task losses with the names and behaviour of the reference's models/loss.py.  SURVEY.md §8 a14:
tiny elementwise losses on (B, out) tensors — they stay PyTorch device ops (not a kernel target);
only the signatures are part of the drop-in surface."""
import torch
import torch.nn.functional as F
from torch import nn


class RMSELoss(nn.Module):
    """sqrt(MSE + eps)  (loss.py:9-16)."""

    def __init__(self, eps=1e-6):
        super().__init__()
        self.mse = nn.MSELoss()
        self.eps = eps

    def forward(self, yhat, y):
        return torch.sqrt(self.mse(yhat, y) + self.eps)


class GHM_Loss(nn.Module):
    """Gradient-harmonised loss base: samples are re-weighted by the (EMA-smoothed) population of
    the gradient-norm bin they fall into (loss.py:19-95)."""

    def __init__(self, bins=10, alpha=0.5):
        super().__init__()
        self._bins = bins
        self._alpha = alpha
        self._last_bin_count = None

    def _g2bin(self, g):
        return torch.floor(g * (self._bins - 0.0001)).long()

    def _custom_loss(self, x, target, weight):
        raise NotImplementedError

    def _custom_loss_grad(self, x, target):
        raise NotImplementedError

    def forward(self, x, target):
        g = self._custom_loss_grad(x, target).abs().detach()
        bin_idx = self._g2bin(g)
        # one histogram instead of `bins` host round trips; same counts (indices outside [0,bins) are ignored)
        inside = (bin_idx >= 0) & (bin_idx < self._bins)
        bin_count = torch.bincount(bin_idx[inside].reshape(-1), minlength=self._bins).float().cpu()
        n = x.size(0) * x.size(1)
        if self._last_bin_count is not None:
            bin_count = self._alpha * self._last_bin_count + (1 - self._alpha) * bin_count
        self._last_bin_count = bin_count
        nonempty = (bin_count > 0).sum().item()
        beta = n / torch.clamp(bin_count * nonempty, min=0.0001)
        beta = beta.to(x.device).type_as(x)
        return self._custom_loss(x, target, beta[bin_idx])


class GHMC_Loss(GHM_Loss):
    def __init__(self, bins, alpha):
        super().__init__(bins, alpha)

    def _custom_loss(self, x, target, weight):
        return F.binary_cross_entropy_with_logits(x, target, weight=weight)

    def _custom_loss_grad(self, x, target):
        return torch.sigmoid(x).detach() - target


class GHMR_Loss(GHM_Loss):
    def __init__(self, bins, alpha, mu):
        super().__init__(bins, alpha)
        self._mu = mu

    def _custom_loss(self, x, target, weight):
        d = x - target
        loss = torch.sqrt(d * d + self._mu * self._mu) - self._mu
        return (loss * weight).sum() / (x.size(0) * x.size(1))

    def _custom_loss_grad(self, x, target):
        d = x - target
        return d / torch.sqrt(d * d + self._mu * self._mu)


class MaskedBCEWithLogitsLoss(nn.Module):
    """BCE-with-logits averaged over the targets that are exactly 0 or 1; NaN targets -> -1 -> masked
    (loss.py:181-200)."""

    def __init__(self):
        super().__init__()
        self.bce_loss = nn.BCEWithLogitsLoss(reduction='none')

    def forward(self, logits, targets):
        targets = torch.where(torch.isnan(targets), torch.full_like(targets, -1), targets)
        mask = (targets == 0) | (targets == 1)
        return (self.bce_loss(logits, targets) * mask).sum() / mask.sum()


def MAEwithNan(y_pred, y_true):
    keep = ~torch.isnan(y_true)
    return F.l1_loss(y_pred[keep], y_true[keep])


def BCEwithNan(y_pred, y_true):
    keep = ~torch.isnan(y_true)
    return F.binary_cross_entropy_with_logits(y_pred[keep], y_true[keep])


def FocalLoss(y_pred, y_true, alpha=0.25, gamma=2):
    """Binary focal loss on probabilities (loss.py:232-254): both classes stacked, p clamped to [1e-5, 1]."""
    if y_pred.shape != y_true.shape:
        y_true = y_true.flatten()
    y_true = y_true.long().float()
    y_pred = y_pred.float()
    t2 = torch.stack((1 - y_true, y_true), dim=1)
    p2 = torch.stack((1 - y_pred, y_pred), dim=1).clamp(1e-5, 1.0)
    loss = -alpha * t2 * torch.pow(1 - p2, gamma) * torch.log(p2)
    return torch.mean(torch.sum(loss, dim=1))


def FocalLossWithLogits(y_pred, y_true, alpha=0.25, gamma=2.0):
    """sigmoid -> keep targets that are exactly 0/1 and not NaN -> FocalLoss with ITS defaults
    (the reference drops alpha/gamma here, loss.py:256-275)."""
    p = torch.sigmoid(y_pred)
    keep = (~torch.isnan(y_true)) & ((y_true == 0.0) | (y_true == 1.0))
    return FocalLoss(p[keep], y_true[keep])


def myCrossEntropyLoss(y_pred, y_true):
    if y_pred.shape != y_true.shape:
        y_true = y_true.flatten()
    return F.cross_entropy(y_pred, y_true)
