"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (pure PyTorch) of the part of the third-party package
``tucker-riemopt == 1.0.1`` that johanDDC/R-TuckER calls.  That package is a
pinned dependency of the reference (poetry.lock:235-252, pyproject.toml:13) but
is NOT vendored in /root/reference and is not installable here (no network), so
this is a restatement of its *published algorithm* (Tucker / SF-Tucker fixed-rank
manifolds: Riemannian gradient by autodiff through the rank-2r tangent
"construct", tangent projection, HOSVD / SF-HOSVD rounding).

PARITY UNPINNED w.r.t. upstream tucker_riemopt: no upstream source, tests or
golden vectors are available.  The restatement is instead pinned by
(i) dense-tensor invariants (tests/test_oracle_manifold.py: construct() equals the
dense tangent vector, project is the orthogonal projector, round() equals the
dense truncated (SF-)HOSVD, grad matches finite differences), and
(ii) the reference's own call sites: with this package on sys.path the
reference's src/model/*/optim.py and train.py:train_one_epoch/evaluate run
UNMODIFIED (tests/test_reference_runs.py).

Import surface that the reference needs:
  train.py:10,197                   set_backend
  train.py:39,41                    Tucker, SFTucker
  asymmetric/optim.py:6-7           Tucker, TuckerRiemannian
  symmetric/optim.py:6-8            SFTucker, SFTuckerRiemannian,
                                    tucker_riemopt.sf_tucker.riemannian.TangentVector
"""
from .tucker.tucker import Tucker
from .sf_tucker.sf_tucker import SFTucker
from .tucker import riemannian as TuckerRiemannian
from .sf_tucker import riemannian as SFTuckerRiemannian

_BACKEND = "pytorch"


def set_backend(name):
    """train.py:197 calls set_backend("pytorch"); only that backend exists here."""
    global _BACKEND
    if name != "pytorch":
        raise ValueError("oracle restatement only implements the pytorch backend")
    _BACKEND = name


__all__ = ["Tucker", "SFTucker", "TuckerRiemannian", "SFTuckerRiemannian", "set_backend"]
