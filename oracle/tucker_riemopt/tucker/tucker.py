"""ORACLE (test infrastructure).  Tucker container restated from the maths.

Call sites in the reference: train.py:41 (construction, factor order
[relation, subject, object]), train.py:79 (``T.norm() ** 2`` inside the loss,
must be differentiable), src/model/asymmetric/optim.py:108 (``.round(rank)``),
:111-114 (``.core`` / ``.factors``).  PARITY UNPINNED vs upstream (see package
docstring).
"""
from dataclasses import dataclass, field
from typing import List, Sequence

import torch


def unfold(t: torch.Tensor, k: int) -> torch.Tensor:
    """Mode-k unfolding (r_k x prod of the others).  Only S_(k) S_(k)^T and left
    singular vectors are ever consumed, so the column order is irrelevant."""
    return torch.movedim(t, k, 0).reshape(t.shape[k], -1)


def mode_dot(t: torch.Tensor, m: torch.Tensor, k: int) -> torch.Tensor:
    """t x_k m  with m of shape (new, old): contracts m's 2nd index with mode k."""
    return torch.movedim(torch.tensordot(m, t, dims=([1], [k])), 0, k)


@dataclass
class Tucker:
    core: torch.Tensor
    factors: List[torch.Tensor] = field(default_factory=list)

    @property
    def rank(self):
        return tuple(self.core.shape)

    @property
    def shape(self):
        return tuple(f.shape[0] for f in self.factors)

    def to_dense(self) -> torch.Tensor:
        t = self.core
        for k, f in enumerate(self.factors):
            t = mode_dot(t, f, k)
        return t

    def flat_inner(self, other: "Tucker") -> torch.Tensor:
        """<self, other>_F through the small r x r' factor Grams."""
        t = self.core
        for k in range(len(self.factors)):
            t = mode_dot(t, other.factors[k].T @ self.factors[k], k)
        return (t * other.core).sum()

    def norm(self) -> torch.Tensor:
        return torch.sqrt(self.flat_inner(self))

    def round(self, rank: Sequence[int]) -> "Tucker":
        """HOSVD truncation to ``rank`` (the retraction, asymmetric/optim.py:108):
        thin QR of every factor, R factors contracted into the core, SVD of every
        unfolding of that small tensor, keep the leading rank[k] left vectors."""
        qs, small = [], self.core
        for k, f in enumerate(self.factors):
            q, r = torch.linalg.qr(f)
            qs.append(q)
            small = mode_dot(small, r, k)
        us = []
        for k in range(len(qs)):
            u, _, _ = torch.linalg.svd(unfold(small, k), full_matrices=False)
            us.append(u[:, : rank[k]])
        core = small
        for k, u in enumerate(us):
            core = mode_dot(core, u.T, k)
        return Tucker(core, [q @ u for q, u in zip(qs, us)])
