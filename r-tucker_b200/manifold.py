"""Plain containers with the names and field order of tucker_riemopt's ``Tucker`` / ``SFTucker``
(reference call sites train.py:39,41): ``extract_tensor`` builds them around the live parameters.
They carry no arithmetic -- the manifold arithmetic lives in the CUDA library."""
from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class Tucker:
    core: torch.Tensor
    factors: List[torch.Tensor] = field(default_factory=list)   # [relation, subject, object]


@dataclass
class SFTucker:
    core: torch.Tensor
    regular_factors: List[torch.Tensor] = field(default_factory=list)   # [relation]
    num_shared_factors: int = 2
    shared_factor: Optional[torch.Tensor] = None                         # entities

    @property
    def factors(self):
        return list(self.regular_factors) + [self.shared_factor] * self.num_shared_factors


def point_tensors(T):
    """(core, R, S, O, sym) of a Tucker-like object."""
    if hasattr(T, "regular_factors"):
        return T.core, T.regular_factors[0], T.shared_factor, T.shared_factor, True
    return T.core, T.factors[0], T.factors[1], T.factors[2], False
