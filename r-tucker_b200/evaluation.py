"""Filtered-ranking evaluation.

Compat functions with the reference's names and semantics on dense tensors:
  filter_predictions   src/utils/utils.py:15-22   (in place, CUDA kernel rt_filter_dense)
  metrics              src/utils/metrics.py:4-22   (sums of 1/rank and hits@k; compare-and-count
                                                    kernel instead of a full sort)
and the fused path ``evaluate`` = train.py:94-125 without dense B x N tensors.
Tie policy: rank = 1 + greater + equal_before (stable descending sort).
"""
import torch

from . import ops
from .manifold import point_tensors


def filter_predictions(predictions, targets, filter):
    f = filter.reshape(-1).to(torch.int32).contiguous()
    return ops.filter_dense_(predictions, targets, f)


def _sums_from_ranks(ranks):
    out = {"mrr": torch.sum(1 / ranks)}
    for k in (1, 3, 10):
        out[f"hits@{k}"] = (ranks <= k).float().sum()
    return out


def metrics(predictions, targets):
    """``targets`` is one-hot after filter_predictions (utils.py:21); the hot column is the target."""
    B = predictions.shape[0]
    tcol = targets.argmax(dim=1).to(torch.int32)
    empty_off = torch.zeros(B + 1, dtype=torch.int32, device=predictions.device)
    empty_idx = torch.zeros(1, dtype=torch.int32, device=predictions.device)
    g, e, eb = ops.rank_filtered(predictions, tcol, empty_off, empty_idx)
    return _sums_from_ranks((1 + g + eb).long())


def rank_batch(model_point, features, filters, n_begin=0, group=None):
    """Fused scoring + ranking of one batch.  features int32 [B,3] = (s, r, o); ``filters`` the CSR of
    all known objects per query.  Returns (ranks int64 [B], bce_sum f64[1], counts)."""
    core, R, S, O, _ = point_tensors(model_point)
    sub, rel, tgt = (features[:, i].contiguous() for i in range(3))
    r_rows = ops.gather_rows(R, rel)
    s_rows = ops.gather_rows(S, sub, n_begin)
    if group is not None:
        torch.distributed.all_reduce(s_rows, group=group)
    q = ops.query_fwd(core, r_rows, s_rows)
    pt = ops.target_prob(q, O, tgt, n_begin)
    if group is not None:
        torch.distributed.all_reduce(pt, group=group)
    g, e, eb, bce = ops.score_rank_fused(q, O, tgt, pt, filters.off, filters.idx, n_begin)
    if group is not None:
        cnt = torch.stack([g, e, eb])
        torch.distributed.all_reduce(cnt, group=group)
        torch.distributed.all_reduce(bce, group=group)
        g, e, eb = cnt[0], cnt[1], cnt[2]
    return (1 + g + eb).long(), bce, (g, e, eb)


@torch.no_grad()
def evaluate(model, dataset, batch_size, device, point=None, n_begin=0, group=None):
    """Fused equivalent of train.py:94-125.  Returns (metrics dict of floats, mean batch BCE)."""
    from .manifold import SFTucker, Tucker
    if point is None:
        if model.symmetric:
            point = SFTucker(model.core.data, [model.R.weight.data], 2, model.E.weight.data)
        else:
            point = Tucker(model.core.data, [model.R.weight.data, model.S.weight.data, model.O.weight.data])
    n_ent = dataset.n_entities
    sums = None
    loss = torch.zeros(1, dtype=torch.float64, device=device)
    nb = 0
    denom = 0
    for features, filters, _, _ in dataset.batches(batch_size, device):
        ranks, bce, _ = rank_batch(point, features, filters, n_begin, group)
        s = _sums_from_ranks(ranks)
        sums = s if sums is None else {k: sums[k] + s[k] for k in s}
        loss += bce / (features.shape[0] * n_ent)     # BCELoss(mean) per batch, train.py:113
        nb += 1
        denom += features.shape[0]
    out = {k: float(v.item()) / denom for k, v in sums.items()}
    return out, loss / nb
