"""Pins the restated tucker_riemopt (oracle/tucker_riemopt) by dense-tensor invariants, since upstream
tucker-riemopt 1.0.1 is not available (parity unpinned w.r.t. upstream; see the package docstring)."""
import pytest
import torch

import tucker_riemopt as tr
from tucker_riemopt import SFTucker, SFTuckerRiemannian as SR, Tucker, TuckerRiemannian as TR
from tucker_riemopt.tucker.tucker import mode_dot, unfold

f64 = torch.float64


def rand_point(sym, g):
    r, n = (3, 4, 4), (7, 9, 9)
    q = lambda a, b: torch.linalg.qr(torch.randn(a, b, generator=g, dtype=f64))[0]
    core = torch.randn(r, generator=g, dtype=f64)
    if sym:
        return SFTucker(core, [q(n[0], r[0])], 2, q(n[1], r[1])), SR
    return Tucker(core, [q(n[i], r[i]) for i in range(3)]), TR


def rand_ambient(sym, g):
    r, n = (2, 3, 3), (7, 9, 9)
    core = torch.randn(r, generator=g, dtype=f64)
    if sym:
        return SFTucker(core, [torch.randn(n[0], r[0], generator=g, dtype=f64)], 2,
                        torch.randn(n[1], r[1], generator=g, dtype=f64))
    return Tucker(core, [torch.randn(n[i], r[i], generator=g, dtype=f64) for i in range(3)])


@pytest.mark.parametrize("sym", [False, True])
def test_projection_is_orthogonal_projector(sym):
    g = torch.Generator().manual_seed(0)
    X, mod = rand_point(sym, g)
    Z = rand_ambient(sym, g)
    tv = mod.project(X, Z)
    d = tv.construct().to_dense()
    assert float((mod.project(X, tv.construct()).construct().to_dense() - d).norm()) < 1e-12    # P∘P = P
    assert abs(float((d * Z.to_dense()).sum() - (d * d).sum())) < 1e-11                            # <PZ,Z> = <PZ,PZ>
    assert abs(float(tv.norm() - d.norm())) < 1e-12 and abs(float(tv.construct().norm() - d.norm())) < 1e-12
    fs = X.regular_factors + [X.shared_factor] if sym else X.factors
    dvs = tv.delta_regular_factors + [tv.delta_shared_factor] if sym else tv.delta_factors
    for u, dv in zip(fs, dvs):
        assert float((u.T @ dv).abs().max()) < 1e-13                                              # gauge


@pytest.mark.parametrize("sym", [False, True])
def test_round_is_truncated_hosvd(sym):
    g = torch.Generator().manual_seed(1)
    X, mod = rand_point(sym, g)
    tv = mod.project(X, rand_ambient(sym, g))
    Y = ((-0.7) * tv + mod.TangentVector(X)).construct()
    r = X.core.shape
    D = Y.to_dense()
    Yr = Y.round(r).to_dense()
    P = D
    if sym:
        u, _, _ = torch.linalg.svd(unfold(D, 0), full_matrices=False)
        P = mode_dot(P, u[:, :r[0]] @ u[:, :r[0]].T, 0)
        ue, _, _ = torch.linalg.svd(torch.cat([unfold(D, 1), unfold(D, 2)], dim=1), full_matrices=False)
        for k in (1, 2):
            P = mode_dot(P, ue[:, :r[1]] @ ue[:, :r[1]].T, k)
    else:
        for k in range(3):
            u, _, _ = torch.linalg.svd(unfold(D, k), full_matrices=False)
            P = mode_dot(P, u[:, :r[k]] @ u[:, :r[k]].T, k)
    assert float((Yr - P).norm() / P.norm()) < 1e-12
    # retraction of the zero step returns the point
    Z0 = mod.TangentVector(X).construct().round(r).to_dense()
    assert float((Z0 - X.to_dense()).norm() / X.to_dense().norm()) < 1e-12


@pytest.mark.parametrize("sym", [False, True])
def test_grad_matches_finite_differences(sym):
    g = torch.Generator().manual_seed(2)
    X, mod = rand_point(sym, g)
    A = torch.randn(X.to_dense().shape, generator=g, dtype=f64)
    f = lambda T: ((T.to_dense() - A) ** 2).sum() + 0.1 * T.norm() ** 2
    rg, fx = mod.grad(f, X)
    eta = mod.project(X, rand_ambient(sym, g))          # a tangent direction
    h = 1e-6
    plus = (h * eta + mod.TangentVector(X)).construct().round(X.core.shape)
    minus = ((-h) * eta + mod.TangentVector(X)).construct().round(X.core.shape)
    fd = float((f(plus) - f(minus)) / (2 * h))
    inner = float((rg.construct().to_dense() * eta.construct().to_dense()).sum())
    assert abs(fd - inner) / abs(inner) < 1e-6
    assert abs(float(fx - f(X))) < 1e-12


def test_set_backend_and_surface():
    tr.set_backend("pytorch")
    with pytest.raises(ValueError):
        tr.set_backend("numpy")
    from tucker_riemopt.sf_tucker.riemannian import TangentVector  # import path used by symmetric/optim.py:8
    assert TangentVector is SR.TangentVector
