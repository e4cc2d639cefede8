"""One call of the variant-3 score op at the WN18RR shape (for ncu launch lists)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtucker_b200 import ops
dev = torch.device("cuda"); B, N, r = 512, 40943, 200
g = torch.Generator().manual_seed(0)
O = torch.linalg.qr(torch.randn(N, r, generator=g))[0].contiguous().to(dev)
q = (3 * torch.randn(B, r, generator=g)).to(dev)
off = torch.arange(0, 2 * B + 1, 2, dtype=torch.int32).to(dev); idx = torch.randint(0, N, (2 * B,), generator=g).int().to(dev)
for _ in range(3):
    ops.score_bce_fwd_bwd(q, q, O, off, idx, 0.1, variant=3)
torch.cuda.synchronize(); print("done")
