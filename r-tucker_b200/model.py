"""R_TuckER models with the reference's class surface.

  asymmetric.R_TuckER   src/model/asymmetric/R_TuckER.py:8-50   (S, R, O embeddings + core)
  symmetric.R_TuckER    src/model/symmetric/R_TuckER.py:8-47    (E, R embeddings + core)

``forward(subject_idx, relation_idx)`` returns a callable ``ScoreFn``; calling it with a
Tucker-like ``T`` returns the dense probabilities [B, n_ent] exactly like the reference's closure
(compat / debug path, computed by the CUDA kernels); the fused training / evaluation paths take the
ScoreFn itself and never materialise B x n_ent.  ``init`` stays in host PyTorch, same RNG call
order as the reference, so identical seeds give identical initial parameters.
"""
import torch
from torch import nn
from torch.nn.init import xavier_normal_, xavier_uniform_

from . import ops
from .manifold import point_tensors


def _i32(t, device):
    return t.to(device=device, dtype=torch.int32).contiguous()


class ScoreFn:
    """What ``model(subject_idx, relation_idx)`` returns (reference: the ``score_fn`` closure)."""

    def __init__(self, model, subject_idx, relation_idx):
        self.model = model
        dev = model.core.device
        self.subject_idx = _i32(subject_idx, dev)
        self.relation_idx = _i32(relation_idx, dev)

    def query(self, T):
        core, R, S, _, _ = point_tensors(T)
        r_rows = ops.gather_rows(R.detach().contiguous(), self.relation_idx)
        s_rows = ops.gather_rows(S.detach().contiguous(), self.subject_idx)
        return ops.query_fwd(core.detach().contiguous(), r_rows, s_rows)

    def __call__(self, T):
        """Dense sigmoid scores [B, n_ent] (R_TuckER.py:42-48)."""
        _, _, _, O, _ = point_tensors(T)
        return ops.score_dense(self.query(T), O.detach().contiguous())


class _Base(nn.Module):
    symmetric = False

    def forward(self, subject_idx, relation_idx):
        return ScoreFn(self, subject_idx, relation_idx)


class AsymmetricRTuckER(_Base):
    symmetric = False

    def __init__(self, data_count, rank=None, **kwargs):
        super().__init__()
        self.S = nn.Embedding(data_count[0], rank[1])
        self.R = nn.Embedding(data_count[1], rank[0])
        self.O = nn.Embedding(data_count[0], rank[2])
        self.core = nn.Parameter(torch.zeros(tuple(rank), dtype=torch.float32))
        self.rank = rank

    def init(self, state_dict=None):
        if state_dict:
            self.load_state_dict(state_dict)
        else:   # same draws, same order as R_TuckER.py:31-39 (on the CPU generator)
            xavier_uniform_(self.core)
            xavier_normal_(self.S.weight)
            xavier_normal_(self.R.weight)
            xavier_normal_(self.O.weight)
            with torch.no_grad():
                self.S.weight.data = torch.linalg.qr(self.S.weight)[0].contiguous()
                self.O.weight.data = torch.linalg.qr(self.O.weight)[0].contiguous()
                self.R.weight.data = torch.linalg.qr(self.R.weight)[0].contiguous()

    def factor_params(self):
        """[R, S, O] in the manifold's mode order (train.py:41)."""
        return [self.R.weight, self.S.weight, self.O.weight]


class SymmetricRTuckER(_Base):
    symmetric = True

    def __init__(self, data_count, rank=None, **kwargs):
        super().__init__()
        self.E = nn.Embedding(data_count[0], rank[1])
        self.R = nn.Embedding(data_count[1], rank[0])
        self.core = nn.Parameter(torch.zeros(tuple(rank), dtype=torch.float32))
        self.rank = rank

    def init(self, state_dict=None):
        if state_dict:
            self.load_state_dict(state_dict)
        else:   # symmetric/R_TuckER.py:29-36
            xavier_uniform_(self.core)
            xavier_normal_(self.E.weight)
            xavier_normal_(self.R.weight)
            with torch.no_grad():
                self.E.weight.data = torch.linalg.qr(self.E.weight)[0].contiguous()
                self.R.weight.data = torch.linalg.qr(self.R.weight)[0].contiguous()

    def factor_params(self):
        return [self.R.weight, self.E.weight]
