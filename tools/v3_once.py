"""Runs the variant-2 score op a few times at the WN18RR shape on bench-like inputs (for ncu)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from rtucker_b200 import ops
dev = torch.device('cuda'); g = torch.Generator().manual_seed(3)
B, N, r2 = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 40943, 200
O = torch.linalg.qr(torch.randn(N, r2, generator=g))[0].contiguous().to(dev)
q = (torch.randn(B, r2, generator=g) * 4 * (N / r2) ** 0.5).to(dev)
off = torch.arange(0, (B + 1) * 2, 2).int().to(dev); idx = torch.randint(0, N, (B * 2,), generator=g).int().to(dev)
for _ in range(4):
    out = ops.score_bce_fwd_bwd(q, None, O, off, idx, 0.1, variant=2, o_absmax=1.0)
torch.cuda.synchronize(); print('done', float(out[0]))
