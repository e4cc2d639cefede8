"""Where does the tensor-core ranking differ from the dense path?  (debug helper for test_score_rank_fused_equals_dense_path)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from rtucker_b200 import ops
from test_gpu_kernels import make_csr
B, N, r2, scale = [float(x) if '.' in x else int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (300, 5000, 64, 30.0)
dev = torch.device("cuda")
g = torch.Generator().manual_seed(7 * B + N + r2)
q = scale * torch.randn(B, r2, generator=g) / r2 ** 0.5
O = torch.randn(N, r2, generator=g)
O[11] = O[3]; O[N - 2] = O[3]; O[N // 2] = O[N // 2 + 1]
target = torch.randint(0, N, (B,), generator=g).int()
target[0], target[1], target[2] = 3, 11, N // 2
z = q @ O.T
top = z.topk(3, dim=1).indices
for b in range(3, B, 2):
    target[b] = int(top[b, b % 3])
off, idx = make_csr(B, N, g, max_per_row=8)
for b in range(B):
    idx[off[b]] = target[b]
qd, Od, td, offd, idxd = (x.to(dev) for x in (q, O, target, off, idx))
P = ops.score_dense(qd, Od)
pt = P[torch.arange(B, device=dev), td.long()].contiguous()
eg, ee, eb = ops.rank_filtered(P.clone(), td, offd, idxd)
cg, ce, cb, bce = ops.score_rank_fused(qd, Od, td, pt, offd, idxd)
bad = ((cg != eg) | (ce != ee) | (cb != eb)).nonzero().flatten().tolist()
print("mismatching queries:", len(bad))
for b in bad[:12]:
    zt = float(z[b, target[b]])
    print(f"b={b} t={int(target[b])} z_t={zt:.4f} p_t={float(pt[b]):.9g} dense (g,e,b)=({int(eg[b])},{int(ee[b])},{int(eb[b])}) "
          f"fused=({int(cg[b])},{int(ce[b])},{int(cb[b])}) filter={idx[off[b]:off[b+1]].tolist()}")
