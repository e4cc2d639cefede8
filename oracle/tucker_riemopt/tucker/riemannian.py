"""ORACLE (test infrastructure).  Riemannian toolbox on the fixed-rank Tucker
manifold, restated from the maths.

Used by the reference as a module: ``TuckerRiemannian.grad / .project /
.TangentVector`` (src/model/asymmetric/optim.py:7,32,52,86-89,107).
PARITY UNPINNED vs upstream tucker_riemopt 1.0.1 (not available).
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from .tucker import Tucker, unfold


def group_cores(corner: torch.Tensor, core: torch.Tensor) -> torch.Tensor:
    """Rank-2r core of a tangent vector: ``corner`` in the [:r,:r,:r] block, ``core``
    in the three blocks that pair one delta-factor with two base factors."""
    r = core.shape
    g = torch.zeros([2 * s for s in r], dtype=core.dtype, device=core.device)
    g[: r[0], : r[1], : r[2]] = corner
    g[r[0]:, : r[1], : r[2]] = core
    g[: r[0], r[1]:, : r[2]] = core
    g[: r[0], : r[1], r[2]:] = core
    return g


@dataclass
class TangentVector:
    point: Tucker
    delta_core: Optional[torch.Tensor] = None
    delta_factors: Optional[List[torch.Tensor]] = None

    def __post_init__(self):
        # TangentVector(X) is the point itself seen as a tangent vector
        # (asymmetric/optim.py:52,107); TangentVector(X, zeros) is the zero vector (:88).
        if self.delta_core is None:
            self.delta_core = self.point.core
        if self.delta_factors is None:
            self.delta_factors = [torch.zeros_like(f) for f in self.point.factors]

    def construct(self) -> Tucker:
        factors = [torch.cat([u, dv], dim=1) for u, dv in zip(self.point.factors, self.delta_factors)]
        return Tucker(group_cores(self.delta_core, self.point.core), factors)

    def __rmul__(self, a):
        return TangentVector(self.point, a * self.delta_core, [a * dv for dv in self.delta_factors])

    def __neg__(self):
        return (-1.0) * self

    def __add__(self, other: "TangentVector"):
        return TangentVector(self.point, self.delta_core + other.delta_core,
                             [a + b for a, b in zip(self.delta_factors, other.delta_factors)])

    def norm(self) -> torch.Tensor:
        """||xi||_F^2 = ||dS||^2 + sum_i tr(dV_i^T dV_i S_(i) S_(i)^T)  (gauge U_i^T dV_i = 0)."""
        core = self.point.core
        sq = (self.delta_core ** 2).sum()
        for k, dv in enumerate(self.delta_factors):
            s = unfold(core, k)
            sq = sq + ((dv.T @ dv) * (s @ s.T)).sum()
        return torch.sqrt(sq)


def _gauge(point: Tucker, d_core: torch.Tensor, d_factors: List[torch.Tensor]) -> TangentVector:
    """Partial derivatives w.r.t. (delta_core, delta_factors) -> tangent vector:
    dV_i <- (I - U_i U_i^T) dV_i (S_(i) S_(i)^T)^-1."""
    dvs = []
    for k, (u, g) in enumerate(zip(point.factors, d_factors)):
        s = unfold(point.core, k)
        g = g - u @ (u.T @ g)
        dvs.append(torch.linalg.solve(s @ s.T, g.T).T)
    return TangentVector(point, d_core, dvs)


def grad(f: Callable[[Tucker], torch.Tensor], x: Tucker):
    """Riemannian gradient of f at x by autodiff through the rank-2r construct,
    evaluated at (delta_core = core, delta_factors = 0).  Returns (TangentVector, f(x))."""
    point = Tucker(x.core.detach(), [u.detach() for u in x.factors])
    dc = point.core.clone().requires_grad_(True)
    dfs = [torch.zeros_like(u).requires_grad_(True) for u in point.factors]
    with torch.enable_grad():
        fx = f(TangentVector(point, dc, dfs).construct())
        grads = torch.autograd.grad(fx, [dc, *dfs])
    return _gauge(point, grads[0], list(grads[1:])), fx.detach()


def project(x: Tucker, xi: Tucker) -> TangentVector:
    """Orthogonal projection of the ambient tensor ``xi`` (any Tucker) onto T_x M:
    Riemannian gradient of  X -> <X, xi>."""
    xi_d = Tucker(xi.core.detach(), [f.detach() for f in xi.factors])
    tv, _ = grad(lambda t: t.flat_inner(xi_d), x)
    return tv
