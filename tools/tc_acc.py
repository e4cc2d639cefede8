"""Accuracy anatomy of the tcgen05 3xTF32 factor update (ops.apply, tc=True) against fp64: norm-wise error, SIGNED mean relative
error (a truncating accumulator shows up as a bias towards zero) and the dependence on the contraction length."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtucker_b200 import ops
dev = torch.device('cuda'); N = 40943
torch.manual_seed(0)
def stats(Y, ref):
    e = (Y.double() - ref)
    nrm = float(e.norm() / ref.norm())
    m = ref.abs() > 0.1 * ref.abs().mean()
    sgn = float(((e / ref)[m]).mean())
    absr = float(((e / ref)[m]).abs().mean())
    return "normwise %.3e  mean signed rel %.3e  mean |rel| %.3e" % (nrm, sgn, absr)
for r in (200,):
    U = torch.linalg.qr(torch.randn(N, r, device=dev))[0].contiguous()
    V = torch.randn(N, r, device=dev); W = torch.randn(N, r, device=dev)
    Y = torch.empty(N, r, device=dev)
    # (1) positive data: every partial sum has the sign and the size of the result -> bias visible
    Up = torch.rand(N, r, device=dev) + 0.5; Kp = (torch.rand(r, r, device=dev, dtype=torch.float64) + 0.5)
    for nk in (1, 2, 3):
        terms = [(Up, Kp)] * nk
        ref = sum(x.double() @ k for x, k in terms)
        for tc in (False, True):
            ops.apply(Y, None, None, terms, tc=tc); print("positive nk=%d tc=%d " % (nk, tc), stats(Y, ref))
    # (2) random signs
    K = torch.randn(r, r, device=dev, dtype=torch.float64)
    ref = V.double() @ K + W.double() @ K
    for tc in (False, True):
        ops.apply(Y, None, None, [(V, K), (W, K)], tc=tc); print("random   nk=2 tc=%d " % tc, stats(Y, ref))
    # (3) near-identity right factor (U Z1 + dV Z2 with Z1 ~ I): the result is reached early, then held
    Z1 = torch.eye(r, device=dev, dtype=torch.float64) + 1e-3 * torch.randn(r, r, device=dev, dtype=torch.float64)
    Z2 = 1e-3 * torch.randn(r, r, device=dev, dtype=torch.float64)
    ref = U.double() @ Z1 + V.double() @ Z2
    for tc in (False, True):
        ops.apply(Y, None, None, [(U, Z1), (V, Z2)], tc=tc); print("near-id  nk=2 tc=%d " % tc, stats(Y, ref))
