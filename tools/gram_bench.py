"""Gram kernel at the WN18RR shape (A^T A: exact symmetric DMMA kernel, csrc/gram_sym.cu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rtucker_b200 import ops
from microbench import timeit
dev = torch.device('cuda'); N, r = 40943, 200
torch.manual_seed(0)
V = torch.randn(N, r, device=dev) * torch.logspace(0, -3, r, device=dev)     # graded columns, like a tangent factor
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = V.double().T @ V.double()
for name, kw in (("default", dict()), ("precise", dict(precise=True))):
    g = ops.gram(V, V, **kw)
    e = (g - ref)
    print("%-8s %.3f ms  normwise %.2e  max rel-to-diag %.2e" % (name, timeit(lambda: ops.gram(V, V, **kw), flush=flush),
          float(e.norm() / ref.norm()), float((e / torch.sqrt(torch.outer(ref.diag(), ref.diag()))).abs().max())))
