"""apply_multi (tcgen05, csrc/apply_tc.cu): accuracy anatomy against fp64 + timing at the WN18RR shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rtucker_b200 import ops
from microbench import timeit
dev = torch.device('cuda'); N, r = 40943, 200
torch.manual_seed(0)
f64 = torch.float64
def stats(Y, ref):
    e = (Y.double() - ref); m = ref.abs() > 0.1 * ref.abs().mean()
    return "normwise %.3e  mean signed rel %.3e" % (float(e.norm() / ref.norm()), float(((e / ref)[m]).mean()))
Up = torch.rand(N, r, device=dev) + 0.5; Kp = torch.rand(r, r, device=dev, dtype=f64) + 0.5
Y = torch.empty(N, r, device=dev)
for nk in (1, 2, 3):
    terms = [(Up, Kp)] * nk
    ref = sum(x.double() @ k for x, k in terms)
    for tc in (False, True):
        ops.apply(Y, None, None, terms, tc=tc); torch.cuda.synchronize(); print("positive nk=%d tc=%d " % (nk, tc), stats(Y, ref))
V = torch.randn(N, r, device=dev); W = torch.randn(N, r, device=dev); K = torch.randn(r, r, device=dev, dtype=f64)
ref = V.double() @ K + W.double() @ K
for tc in (False, True):
    ops.apply(Y, None, None, [(V, K), (W, K)], tc=tc); print("random   nk=2 tc=%d " % tc, stats(Y, ref))
# two jobs, in place on a term, copies of the operands, scalar on X0
U1 = torch.linalg.qr(torch.randn(N, r, device=dev))[0].contiguous(); U2 = U1.clone() + 0.01
Z1 = torch.eye(r, device=dev, dtype=f64) + 1e-3 * torch.randn(r, r, device=dev, dtype=f64); Z2 = 1e-3 * torch.randn(r, r, device=dev, dtype=f64)
a0 = torch.tensor([0.37], dtype=f64, device=dev)
ref1 = U1.double() @ Z1 + V.double() @ Z2
ref2 = 0.37 * W.double() + U2.double() @ Z1 + V.double() @ Z2 + W.double() @ K
U1c, Vc = torch.zeros_like(U1), torch.zeros_like(V)
U1w = U1.clone(); Y2 = torch.empty(N, r, device=dev)
ops.apply_multi([(U1w, None, None, [(U1w, Z1, U1c), (V, Z2, Vc)]), (Y2, W, a0, [(U2, Z1), (V, Z2), (W, K)])])
torch.cuda.synchronize()
print("job1 in place ", stats(U1w, ref1), " copies exact:", bool(torch.equal(U1c, U1)), bool(torch.equal(Vc, V)))
print("job2 3 terms  ", stats(Y2, ref2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for tc in (False, True):
    print('tc', tc, 'apply1 %.3f ms' % timeit(lambda: ops.apply(Y, Y, None, [(V, K)], tc=tc), flush=flush),
          'apply2 %.3f' % timeit(lambda: ops.apply(Y, None, None, [(V, K), (W, K)], tc=tc), flush=flush),
          'apply3 %.3f' % timeit(lambda: ops.apply(Y, Y, None, [(U1, K), (V, K), (W, K)], tc=tc), flush=flush))
print('multi: 2 jobs x 3 terms %.3f ms' % timeit(lambda: ops.apply_multi([(Y, Y, a0, [(U1, K), (V, K), (W, K)]), (Y2, Y2, a0, [(U2, K), (V, K), (W, K)])]), flush=flush))
print('multi: retraction-like 2 jobs x 2 terms with copies %.3f ms' % timeit(lambda: ops.apply_multi([(U1w, None, None, [(U1w, Z1, U1c), (V, Z2, Vc)]), (U2, None, None, [(U2, Z1, Y2), (W, Z2, Y)])]), flush=flush))
