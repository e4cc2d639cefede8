"""Role cycle counters of the rank-mode / score-mode tcgen05 kernel (profiling build, see tools/apply_prof.py)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtucker_b200 import ops
dev = torch.device("cuda"); B, N, r = 512, 40943, 200
g = torch.Generator().manual_seed(0)
O = torch.linalg.qr(torch.randn(N, r, generator=g))[0].contiguous().to(dev)
off = torch.arange(0, 16 * B + 1, 16, dtype=torch.int32).to(dev); idx = torch.randint(0, N, (16 * B,), generator=g).int().to(dev)
tgt = idx[::16].contiguous()
q = (3.0 * (N / r) ** 0.5 * torch.randn(B, r, generator=g) / r ** 0.5).to(dev)
pt = ops.target_prob(q, O, tgt)
print("== rank"); ops.score_rank_fused(q, O, tgt, pt, off, idx); torch.cuda.synchronize()
print("== score"); ops.score_bce_fwd_bwd(q, q, O, off, idx, 0.1, variant=3); torch.cuda.synchronize()
