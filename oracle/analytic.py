"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Closed-form ("analytic") CPU restatement of one R-TuckER training step, stage by
stage, at rank r (no autodiff, no rank-2r construct).  This is the executable
specification the CUDA kernels are checked against one stage at a time; it is
itself pinned against (i) the reference's own ``R_TuckER.forward`` + torch
autograd and (ii) the upstream-style restatement in ``oracle/tucker_riemopt``
driving the reference's unmodified optimisers (tests/test_oracle_analytic.py).

Reference lines each stage follows:
  query_fwd           src/model/asymmetric/R_TuckER.py:43-46 (einsum + bmm)
  score_bce_fwd_bwd   src/model/asymmetric/R_TuckER.py:47-48, train.py:79,136 (nn.BCELoss mean),
                      src/data/Dataset.py:50-52 (label smoothing)
  riemannian_grad     tucker_riemopt.grad as called at asymmetric/optim.py:89, symmetric/optim.py:83
  tangent_norm        asymmetric/optim.py:90
  project             asymmetric/optim.py:86, symmetric/optim.py:80
  retract             asymmetric/optim.py:106-109 (construct().round(rank))
  RSGDState.fit/step  asymmetric/optim.py:74-114, symmetric/optim.py:23-107
  AdamState.fit/step  symmetric/optim.py:110-167 (SFTuckerAdam)

Conventions: modes are (relation, subject, object) (train.py:37-42); ``sym=True``
means SF-Tucker with subject and object modes sharing the factor ``E``.
All tensors are torch CPU tensors of one dtype (float64 for the checker, float32
to mimic the reference's arithmetic).
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

BCE_LOG_CLAMP = -100.0   # torch.nn.BCELoss clamps log terms at -100
BCE_GRAD_EPS = 1e-12     # binary_cross_entropy_backward: max((1-p)*p, 1e-12)


def unfold(t, k):
    return torch.movedim(t, k, 0).reshape(t.shape[k], -1)


def mode_dot(t, m, k):
    return torch.movedim(torch.tensordot(m, t, dims=([1], [k])), 0, k)


def multi_mode_dot(t, ms, skip=None):
    for k, m in enumerate(ms):
        if k != skip and m is not None:
            t = mode_dot(t, m, k)
    return t


# ----------------------------------------------------------------------------------------------
# (a) query contraction and its backward
# ----------------------------------------------------------------------------------------------
def query_fwd(core, R, S, rel_idx, sub_idx):
    """q[b,:] = core x_1 R[rel_b] x_2 S[sub_b]     (asymmetric/R_TuckER.py:43-46)."""
    return torch.einsum("aij,ba,bi->bj", core, R[rel_idx], S[sub_idx])


def query_bwd(core, R, S, rel_idx, sub_idx, H):
    """Given H = dL/dq [B,r_O]: dCore [r_R,r_S,r_O], per-query rows dS_rows [B,r_S], dR_rows [B,r_R]."""
    r, s = R[rel_idx], S[sub_idx]
    d_core = torch.einsum("ba,bi,bj->aij", r, s, H)
    y = torch.einsum("bj,aij->bai", H, core)
    ds_rows = torch.einsum("ba,bai->bi", r, y)
    dr_rows = torch.einsum("bi,bai->ba", s, y)
    return d_core, ds_rows, dr_rows


# ----------------------------------------------------------------------------------------------
# (b) 1-N score + sigmoid + label-smoothed BCE (mean) + backward
# ----------------------------------------------------------------------------------------------
def dense_targets(B, N, tgt_off, tgt_idx, label_smoothing, dtype):
    """Multi-hot targets as KG_dataset builds them (src/data/Dataset.py:43-52)."""
    t = torch.zeros(B, N, dtype=dtype)
    for b in range(B):
        t[b, tgt_idx[tgt_off[b]:tgt_off[b + 1]].long()] = 1
    if label_smoothing > 0:
        t = (1 - label_smoothing) * t + label_smoothing / N
    return t


# When True, the elementwise sigmoid / log / gradient terms are evaluated in float32 even if the
# contractions run in float64: this reproduces the reference's fp32 saturation semantics (p rounds to
# exactly 1.0 for z > 16.6, log terms clamp at -100, the gradient vanishes where p(1-p) < 1e-12,
# SURVEY.md App. B.1) while keeping the linear algebra exact.
ELEMENTWISE_FP32 = False


def bce_sigmoid_terms(z, t):
    """Elementwise loss and dL_sum/dz exactly as BCELoss(sigmoid(z)) computes them
    (before the 1/(B*N) mean factor)."""
    if ELEMENTWISE_FP32 and z.dtype != torch.float32:
        p, loss, g = bce_sigmoid_terms(z.float(), t.float())
        return p.to(z.dtype), loss.to(z.dtype), g.to(z.dtype)
    p = torch.sigmoid(z)
    loss = -(t * torch.clamp(torch.log(p), min=BCE_LOG_CLAMP)
             + (1 - t) * torch.clamp(torch.log1p(-p), min=BCE_LOG_CLAMP))
    pq = (1 - p) * p
    g = (p - t) / torch.clamp(pq, min=BCE_GRAD_EPS) * pq
    return p, loss, g


def score_bce_fwd_bwd(q, O, tgt_off, tgt_idx, label_smoothing, n_total=None, batch_total=None):
    """loss = mean BCE(sigmoid(q O^T), T);  H = G O;  dO = G^T q  with G = dloss/dZ.
    ``n_total`` / ``batch_total`` give the mean's denominator when O is an entity
    shard (entity-sharded path: the loss and H returned are then partial sums)."""
    B, N = q.shape[0], O.shape[0]
    n_total = N if n_total is None else n_total
    batch_total = B if batch_total is None else batch_total
    t = torch.zeros(B, N, dtype=q.dtype)
    for b in range(B):
        ids = tgt_idx[tgt_off[b]:tgt_off[b + 1]].long()
        t[b, ids] = 1
    if label_smoothing > 0:
        t = (1 - label_smoothing) * t + label_smoothing / n_total
    z = q @ O.T
    _, loss, g = bce_sigmoid_terms(z, t)
    inv = 1.0 / (batch_total * n_total)
    G = g * inv
    return loss.sum() * inv, G @ O, G.T @ q


# ----------------------------------------------------------------------------------------------
# (c) manifold pieces
# ----------------------------------------------------------------------------------------------
@dataclass
class Point:
    """A point of the (SF-)Tucker manifold.  factors = [R, S, O]; when sym, S is O is E."""
    core: torch.Tensor
    factors: List[torch.Tensor]
    sym: bool = False

    def gram_inv_list(self):
        """(S_(i) S_(i)^T) per mode; for SF the shared modes carry the SUM of both."""
        g = [unfold(self.core, k) @ unfold(self.core, k).T for k in range(3)]
        if self.sym:
            g[1] = g[1] + g[2]
            g[2] = g[1]
        return g

    def to_dense(self):
        return multi_mode_dot(self.core, self.factors)


@dataclass
class Tangent:
    """(dS; dV_R, dV_S, dV_O).  For sym, dV_S is dV_O is dV_E (stored twice as the same object)."""
    d_core: torch.Tensor
    d_factors: List[torch.Tensor]

    def to_dense(self, x: Point):
        out = multi_mode_dot(self.d_core, x.factors)
        for k in range(3):
            fs = list(x.factors)
            fs[k] = self.d_factors[k]
            out = out + multi_mode_dot(x.core, fs)
        return out


def scatter_rows(n, idx, rows):
    out = torch.zeros(n, rows.shape[1], dtype=rows.dtype)
    out.index_add_(0, idx.long(), rows)
    return out


def riemannian_grad(x: Point, rel_idx, sub_idx, tgt_off, tgt_idx, label_smoothing, reg):
    """Closed form of tucker_riemopt.grad for the reference loss.
    Returns (Tangent, loss, intermediates dict)."""
    R, S, O = x.factors
    q = query_fwd(x.core, R, S, rel_idx, sub_idx)
    bce, H, dO = score_bce_fwd_bwd(q, O, tgt_off, tgt_idx, label_smoothing)
    d_core, ds_rows, dr_rows = query_bwd(x.core, R, S, rel_idx, sub_idx, H)
    loss = bce + reg * (x.core ** 2).sum()
    dS = d_core + 2 * reg * x.core
    gR = scatter_rows(R.shape[0], rel_idx, dr_rows)
    gS = scatter_rows(S.shape[0], sub_idx, ds_rows)
    gO = dO
    grams = x.gram_inv_list()
    if x.sym:
        gE = gS + gO
        raw = [gR, gE, gE]
    else:
        raw = [gR, gS, gO]
    dvs = []
    for k in range(3):
        if x.sym and k == 2:
            dvs.append(dvs[1])
            continue
        u = x.factors[k]
        g = raw[k] - u @ (u.T @ raw[k])
        dvs.append(torch.linalg.solve(grams[k], g.T).T)
    inter = dict(q=q, bce=bce, H=H, dO=dO, d_core=d_core, ds_rows=ds_rows, dr_rows=dr_rows)
    return Tangent(dS, dvs), loss, inter


def tangent_norm(x: Point, xi: Tangent):
    grams = x.gram_inv_list()
    sq = (xi.d_core ** 2).sum()
    for k in range(3):
        if x.sym and k == 2:
            continue
        dv = xi.d_factors[k]
        sq = sq + ((dv.T @ dv) * grams[k]).sum()
    return torch.sqrt(sq)


def group_cores(corner, core):
    r = core.shape
    g = torch.zeros([2 * s for s in r], dtype=core.dtype)
    g[: r[0], : r[1], : r[2]] = corner
    g[r[0]:, : r[1], : r[2]] = core
    g[: r[0], r[1]:, : r[2]] = core
    g[: r[0], : r[1], r[2]:] = core
    return g


def project(x: Point, old: Point, xi_old: Tangent):
    """Projection onto T_x of the ambient tensor xi_old (a tangent vector at ``old``),
    i.e. the vector transport of asymmetric/optim.py:86."""
    G = group_cores(xi_old.d_core, old.core)
    F = [torch.cat([u, dv], dim=1) for u, dv in zip(old.factors, xi_old.d_factors)]
    M = [x.factors[k].T @ F[k] for k in range(3)]           # r_k x 2r_k
    pS = multi_mode_dot(G, M)
    grams = x.gram_inv_list()
    KC = []                                                  # Y_k(k) C_(k)^T  (2r_k x r_k)
    for k in range(3):
        Yk = multi_mode_dot(G, M, skip=k)
        KC.append(unfold(Yk, k) @ unfold(x.core, k).T)
    if x.sym:
        KC[1] = KC[1] + KC[2]
        KC[2] = KC[1]
    pV = []
    for k in range(3):
        if x.sym and k == 2:
            pV.append(pV[1])
            continue
        K = torch.linalg.solve(grams[k], KC[k].T).T          # (2r x r) @ inv(gram)
        pV.append(F[k] @ K - x.factors[k] @ (M[k] @ K))
    return Tangent(pS, pV)


def retract(x: Point, xi: Tangent, lr, rank=None):
    """round(construct(X - lr*xi)) by the structured route of SURVEY.md App. A.5:
    thin QR of W_i = -lr dV_i (U_i^T W_i = 0, so [U_i|W_i] = [U_i|Q_i] blkdiag(I,R_i)), small rank-2r
    tensor, HOSVD through the symmetric eigenproblem of each unfolding Gram.  Returns the new Point."""
    r = x.core.shape if rank is None else rank
    Rf, W, Q = [], [], []
    for k in range(3):
        w = -lr * xi.d_factors[k]
        q, rr = torch.linalg.qr(w)       # Householder QR: fine for rank-deficient W (e.g. M - r0 < r0)
        Q.append(q)
        Rf.append(rr)
        W.append(w)
    T = group_cores(x.core - lr * xi.d_core, x.core)
    blk = []
    for k in range(3):
        b = torch.zeros(2 * r[k], 2 * r[k], dtype=T.dtype)
        b[: r[k], : r[k]] = torch.eye(r[k], dtype=T.dtype)
        b[r[k]:, r[k]:] = Rf[k]
        blk.append(b)
    T = multi_mode_dot(T, blk)
    Ng = [unfold(T, k) @ unfold(T, k).T for k in range(3)]
    if x.sym:
        Ng[1] = Ng[1] + Ng[2]
        Ng[2] = Ng[1]
    Y = []
    for k in range(3):
        if x.sym and k == 2:
            Y.append(Y[1])
            continue
        evals, evecs = torch.linalg.eigh(Ng[k])
        Y.append(evecs[:, torch.argsort(evals, descending=True)[: r[k]]])
    new_core = multi_mode_dot(T, [y.T for y in Y])
    new_f = []
    for k in range(3):
        if x.sym and k == 2:
            new_f.append(new_f[1])
            continue
        y1, y2 = Y[k][: r[k]], Y[k][r[k]:]
        new_f.append(x.factors[k] @ y1 + Q[k] @ y2)
    return Point(new_core, new_f, x.sym)


def axpby(a, xi: Tangent, b, eta: Optional[Tangent]):
    if eta is None:
        dvs = [a * v for v in xi.d_factors]
        return Tangent(a * xi.d_core, dvs)
    dvs = [a * v + b * w for v, w in zip(xi.d_factors, eta.d_factors)]
    return Tangent(a * xi.d_core + b * eta.d_core, dvs)


class RSGDState:
    """Analytic twin of RSGDwithMomentum (momentum_beta>0) / RGD (momentum_beta=None):
    asymmetric/optim.py:74-114, symmetric/optim.py:23-107."""

    def __init__(self, x: Point, momentum_beta: Optional[float] = 0.8):
        self.x = x
        self.beta = momentum_beta
        self.old: Optional[Point] = None
        self.direction: Optional[Tangent] = None
        self.loss = None

    def fit(self, rel_idx, sub_idx, tgt_off, tgt_idx, label_smoothing, reg, normalize_grad=1.0):
        momentum = None
        if self.beta is not None and self.direction is not None:
            momentum = project(self.x, self.old, self.direction)
        g, self.loss, self.inter = riemannian_grad(self.x, rel_idx, sub_idx, tgt_off, tgt_idx,
                                                   label_smoothing, reg)
        norm = tangent_norm(self.x, g)
        scale = 1.0 if not normalize_grad else normalize_grad / norm
        self.rgrad = g
        self.momentum = momentum
        self.direction = axpby(scale, g, self.beta, momentum)
        return norm

    def step(self, lr):
        new = retract(self.x, self.direction, lr)
        self.old = self.x
        self.x = new
        return new


class AdamState:
    """Analytic twin of SFTuckerAdam (symmetric/optim.py:110-167); the same update on the Tucker manifold when
    ``x.sym`` is False.  The gradient is NOT normalised (fit's normalize_grad argument is unused upstream).

    ``live_point_quirk``.  In the reference the stored momentum is a TangentVector whose ``point`` wraps the LIVE
    parameters (train.py:37-42 builds the point around ``param.data`` without copying) and ``momentum.construct()``
    is only evaluated in the NEXT fit (optim.py:136), after step() has overwritten the parameters in place
    (optim.py:163-165).  The ambient tensor that gets projected is then built from the OLD deltas (dS, dV) and the
    NEW core / factors -- unlike RSGDwithMomentum, which constructs before the write-back (asymmetric/optim.py:109).
    That tensor depends on the GAUGE of the new point (sign / rotation conventions of the SVD inside ``round``), so
    the reference's Adam trajectory is not a function of the tensors alone and cannot be matched by any retraction
    with another (equivalent) gauge.  ``live_point_quirk=True`` reproduces it exactly when the twin is re-seeded with
    the reference's own (core, factors) before every step (tests/test_oracle_analytic.py::
    test_live_reference_sftucker_adam); the default ``False`` transports the momentum as the ambient tensor at the
    point where it was formed -- the gauge-invariant update the CUDA engine implements."""

    def __init__(self, x: Point, betas=(0.9, 0.999), eps=1e-8, step_velocity=1, live_point_quirk=False):
        self.x = x
        self.betas, self.eps, self.step_velocity = betas, eps, step_velocity
        self.live_point_quirk = live_point_quirk
        self.momentum: Optional[Tangent] = None
        self.momentum_point: Optional[Point] = None
        self.second_momentum = 0.0
        self.step_t = 1
        self.loss = None

    def fit(self, rel_idx, sub_idx, tgt_off, tgt_idx, label_smoothing, reg):
        b1, b2 = self.betas
        g, self.loss, self.inter = riemannian_grad(self.x, rel_idx, sub_idx, tgt_off, tgt_idx, label_smoothing, reg)
        norm = tangent_norm(self.x, g)
        if self.momentum is not None:                                   # optim.py:135-137
            base = self.x if self.live_point_quirk else self.momentum_point
            transported = project(self.x, base, self.momentum)
            self.momentum = axpby(b1, transported, 1 - b1, g)
        else:                                                           # optim.py:139
            self.momentum = axpby(1 - b1, g, 0.0, None)
        self.momentum_point = self.x
        self.second_momentum = b2 * self.second_momentum + (1 - b2) * float(norm) ** 2          # optim.py:140
        e = self.step_t // self.step_velocity + 1
        corrected = self.second_momentum / (1 - b2 ** e)                                        # optim.py:141
        ratio = (1 - b1 ** e) * corrected ** 0.5 + self.eps                                     # optim.py:142-144
        self.ratio = ratio
        self.direction = axpby(1.0 / ratio, self.momentum, 0.0, None)                           # optim.py:145
        return norm

    def step(self, lr):
        new = retract(self.x, self.direction, lr)                                               # optim.py:157-161
        self.x = new
        self.step_t += 1                                                                        # optim.py:167
        return new
