"""ORACLE / TEST INFRASTRUCTURE + CPU BASELINE ("port") -- never imported by the product path.

A line-by-line PORT (not a copy) of what the reference executes for one training step and one
evaluation batch on the CPU, written against the restated toolbox in oracle/tucker_riemopt, so it can
run where /root/reference does not exist (the GPU box):

  score_fn            src/model/asymmetric/R_TuckER.py:41-50 / symmetric/R_TuckER.py:38-47
  loss closure        train.py:79   BCELoss(mean)(score_fn(T), dense targets) + reg * T.norm()**2
  fit / step          src/model/asymmetric/optim.py:74-114, symmetric/optim.py:23-107
                      (momentum transport by project, Riemannian gradient by autodiff through the
                      rank-2r construct, unit-normalised direction, HOSVD retraction)
  dense targets       src/data/Dataset.py:43-52
  eval batch          train.py:107-117 + src/utils/utils.py:15-22 + src/utils/metrics.py:4-22

tests/test_oracle_reference_step.py checks it against the reference's own classes where the
checkout exists.  bench.py times it as ``cpu_baseline`` (kind "port") and as ``--impl reference``.
"""
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from tucker_riemopt import SFTucker, SFTuckerRiemannian, Tucker, TuckerRiemannian  # noqa: E402


def dense_targets(n_ent, off, idx, label_smoothing):
    B = off.shape[0] - 1
    t = torch.zeros(B, n_ent)
    rows = torch.repeat_interleave(torch.arange(B), (off[1:] - off[:-1]).long())
    t[rows, idx.long()] = 1
    if label_smoothing > 0:
        t = (1 - label_smoothing) * t + label_smoothing / n_ent
    return t


def make_score_fn(sym, subject_idx, relation_idx):
    def score_fn(T):
        if sym:
            relations = T.regular_factors[0][relation_idx, :]
            subjects = T.shared_factor[subject_idx, :]
            objects = T.shared_factor
        else:
            relations = T.factors[0][relation_idx, :]
            subjects = T.factors[1][subject_idx, :]
            objects = T.factors[2]
        preds = torch.einsum("abc,da->dbc", T.core, relations)
        preds = torch.bmm(subjects.view(-1, 1, subjects.shape[1]), preds).view(-1, subjects.shape[1])
        preds = preds @ objects.T
        return torch.sigmoid(preds)
    return score_fn


class ReferenceStepper:
    """State of the reference's RSGDwithMomentum (momentum_beta given) or RGD (None)."""

    def __init__(self, core, R, S, O=None, momentum_beta=0.8):
        self.sym = O is None
        self.core, self.R, self.S, self.O = core, R, S, O
        self.beta = momentum_beta
        self.direction = None
        self.loss = None
        self.mod = SFTuckerRiemannian if self.sym else TuckerRiemannian
        self.rank = tuple(core.shape)

    def point(self):
        if self.sym:
            return SFTucker(self.core, [self.R], 2, self.S)
        return Tucker(self.core, [self.R, self.S, self.O])

    def train_step(self, subject_idx, relation_idx, targets, reg, lr, normalize_grad=1.0):
        criterion = torch.nn.BCELoss(reduction="mean")
        score_fn = make_score_fn(self.sym, subject_idx, relation_idx)
        loss_fn = lambda T: criterion(score_fn(T), targets) + reg * T.norm() ** 2  # noqa: E731
        x_k = self.point()
        mod = self.mod
        if self.beta is not None:
            if self.direction is not None:
                momentum = mod.project(x_k, self.direction)
            else:
                momentum = mod.TangentVector(x_k, torch.zeros_like(x_k.core))
        rgrad, self.loss = mod.grad(loss_fn, x_k)
        rgrad_norm = rgrad.norm().detach()
        normalize = rgrad_norm if not normalize_grad else normalize_grad
        direction = (1 / rgrad_norm * normalize) * rgrad
        if self.beta is not None:
            direction = direction + self.beta * momentum
        with torch.no_grad():
            x_new = ((-lr) * direction + mod.TangentVector(direction.point)).construct().round(self.rank)
            if self.beta is not None:
                self.direction = direction.construct()
        self.core = x_new.core
        if self.sym:
            self.R, self.S = x_new.regular_factors[0], x_new.shared_factor
        else:
            self.R, self.S, self.O = x_new.factors
        return rgrad_norm

    @torch.no_grad()
    def eval_batch(self, features, targets):
        """features [B,3] (s, r, o); targets dense multi-hot.  Returns (metric sums, batch BCE)."""
        score_fn = make_score_fn(self.sym, features[:, 0], features[:, 1])
        predictions = score_fn(self.point())
        loss = torch.nn.BCELoss(reduction="mean")(predictions, targets)
        flt = features[:, 2].reshape(-1, 1)
        keep = predictions.gather(1, flt)
        predictions[targets == 1] = 0
        targets[targets == 1] = 0
        predictions.scatter_(1, flt, keep)
        targets.scatter_(1, flt, torch.ones(keep.shape))
        _, idx = torch.sort(predictions, dim=1, descending=True)
        ts = targets.gather(1, idx)
        ranks = ts.argmax(dim=1) + 1
        out = {"mrr": torch.sum(1 / ranks)}
        for k in (1, 3, 10):
            h = ts[:, :k].sum(dim=1).float()
            h[h > 1] = 1
            out[f"hits@{k}"] = h.sum()
        return out, loss, ranks
