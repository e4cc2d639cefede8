"""ORACLE (test infrastructure).  Riemannian toolbox on the SF-Tucker manifold
(shared factor in the last two modes).

Reference call sites: src/model/symmetric/optim.py:7-8,34-39,54,80-86,101,133-145,160
(``SFTuckerRiemannian.grad/.project/.TangentVector`` and the by-path import of
``TangentVector``).  PARITY UNPINNED vs upstream tucker_riemopt 1.0.1.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from ..tucker.riemannian import group_cores
from ..tucker.tucker import unfold
from .sf_tucker import SFTucker


@dataclass
class TangentVector:
    point: SFTucker
    delta_core: Optional[torch.Tensor] = None
    delta_regular_factors: Optional[List[torch.Tensor]] = None
    delta_shared_factor: Optional[torch.Tensor] = None

    def __post_init__(self):
        if self.delta_core is None:
            self.delta_core = self.point.core
        if self.delta_regular_factors is None:
            self.delta_regular_factors = [torch.zeros_like(f) for f in self.point.regular_factors]
        if self.delta_shared_factor is None:
            self.delta_shared_factor = torch.zeros_like(self.point.shared_factor)

    def construct(self) -> SFTucker:
        p = self.point
        regular = [torch.cat([u, dv], dim=1) for u, dv in zip(p.regular_factors, self.delta_regular_factors)]
        shared = torch.cat([p.shared_factor, self.delta_shared_factor], dim=1)
        return SFTucker(group_cores(self.delta_core, p.core), regular, p.num_shared_factors, shared)

    def __rmul__(self, a):
        return TangentVector(self.point, a * self.delta_core,
                             [a * dv for dv in self.delta_regular_factors], a * self.delta_shared_factor)

    def __neg__(self):
        return (-1.0) * self

    def __add__(self, other: "TangentVector"):
        return TangentVector(self.point, self.delta_core + other.delta_core,
                             [a + b for a, b in zip(self.delta_regular_factors, other.delta_regular_factors)],
                             self.delta_shared_factor + other.delta_shared_factor)

    def norm(self) -> torch.Tensor:
        p = self.point
        nreg = len(p.regular_factors)
        sq = (self.delta_core ** 2).sum()
        for k, dv in enumerate(self.delta_regular_factors):
            s = unfold(p.core, k)
            sq = sq + ((dv.T @ dv) * (s @ s.T)).sum()
        gs = sum(unfold(p.core, k) @ unfold(p.core, k).T for k in range(nreg, nreg + p.num_shared_factors))
        de = self.delta_shared_factor
        sq = sq + ((de.T @ de) * gs).sum()
        return torch.sqrt(sq)


def _gauge(point: SFTucker, d_core, d_regular, d_shared) -> TangentVector:
    nreg = len(point.regular_factors)
    dvs = []
    for k, (u, g) in enumerate(zip(point.regular_factors, d_regular)):
        s = unfold(point.core, k)
        g = g - u @ (u.T @ g)
        dvs.append(torch.linalg.solve(s @ s.T, g.T).T)
    e = point.shared_factor
    gs = sum(unfold(point.core, k) @ unfold(point.core, k).T for k in range(nreg, nreg + point.num_shared_factors))
    g = d_shared - e @ (e.T @ d_shared)
    de = torch.linalg.solve(gs, g.T).T
    return TangentVector(point, d_core, dvs, de)


def grad(f: Callable[[SFTucker], torch.Tensor], x: SFTucker):
    point = SFTucker(x.core.detach(), [u.detach() for u in x.regular_factors],
                     x.num_shared_factors, x.shared_factor.detach())
    dc = point.core.clone().requires_grad_(True)
    drs = [torch.zeros_like(u).requires_grad_(True) for u in point.regular_factors]
    de = torch.zeros_like(point.shared_factor).requires_grad_(True)
    with torch.enable_grad():
        fx = f(TangentVector(point, dc, drs, de).construct())
        grads = torch.autograd.grad(fx, [dc, *drs, de])
    return _gauge(point, grads[0], list(grads[1:-1]), grads[-1]), fx.detach()


def project(x: SFTucker, xi: SFTucker) -> TangentVector:
    xi_d = SFTucker(xi.core.detach(), [f.detach() for f in xi.regular_factors],
                    xi.num_shared_factors, xi.shared_factor.detach())
    tv, _ = grad(lambda t: t.flat_inner(xi_d), x)
    return tv
