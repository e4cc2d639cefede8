"""ORACLE (test infrastructure).  Shared-factor Tucker container restated from the maths.

Reference call sites: train.py:39 (``SFTucker(core, [R], num_shared_factors=2,
shared_factor=E)``), train.py:79 (``.norm()``), src/model/symmetric/optim.py:55-59
(``.round(rank)``, ``.core``, ``.regular_factors``, ``.shared_factor``).  The shared
modes are the LAST ``num_shared_factors`` modes.  PARITY UNPINNED vs upstream.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from ..tucker.tucker import mode_dot, unfold


@dataclass
class SFTucker:
    core: torch.Tensor
    regular_factors: List[torch.Tensor] = field(default_factory=list)
    num_shared_factors: int = 2
    shared_factor: Optional[torch.Tensor] = None

    @property
    def factors(self) -> List[torch.Tensor]:
        return list(self.regular_factors) + [self.shared_factor] * self.num_shared_factors

    @property
    def rank(self):
        return tuple(self.core.shape)

    def to_dense(self) -> torch.Tensor:
        t = self.core
        for k, f in enumerate(self.factors):
            t = mode_dot(t, f, k)
        return t

    def flat_inner(self, other: "SFTucker") -> torch.Tensor:
        t = self.core
        for k, (fa, fb) in enumerate(zip(self.factors, other.factors)):
            t = mode_dot(t, fb.T @ fa, k)
        return (t * other.core).sum()

    def norm(self) -> torch.Tensor:
        return torch.sqrt(self.flat_inner(self))

    def round(self, rank: Sequence[int]) -> "SFTucker":
        """SF-HOSVD truncation: regular modes as in HOSVD; the shared factor takes the
        leading left singular vectors of the concatenated shared-mode unfoldings."""
        nreg = len(self.regular_factors)
        small, qs = self.core, []
        for k, f in enumerate(self.regular_factors):
            q, r = torch.linalg.qr(f)
            qs.append(q)
            small = mode_dot(small, r, k)
        qe, re = torch.linalg.qr(self.shared_factor)
        for k in range(nreg, nreg + self.num_shared_factors):
            small = mode_dot(small, re, k)
        us = []
        for k in range(nreg):
            u, _, _ = torch.linalg.svd(unfold(small, k), full_matrices=False)
            us.append(u[:, : rank[k]])
        cat = torch.cat([unfold(small, k) for k in range(nreg, nreg + self.num_shared_factors)], dim=1)
        ue, _, _ = torch.linalg.svd(cat, full_matrices=False)
        ue = ue[:, : rank[nreg]]
        core = small
        for k, u in enumerate(us):
            core = mode_dot(core, u.T, k)
        for k in range(nreg, nreg + self.num_shared_factors):
            core = mode_dot(core, ue.T, k)
        return SFTucker(core, [q @ u for q, u in zip(qs, us)], self.num_shared_factors, qe @ ue)
