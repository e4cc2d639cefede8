"""One call of the HOSVD subspace kernel on a graded 400 x 400 matrix (ncu target)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from rtucker_b200 import ops  # noqa: E402
from subspace_bench import graded  # noqa: E402
A = graded(400, 200, 1e6, 0.3, 1.3, 1).to("cuda:0")
for _ in range(2):
    Y, info = ops.dominant_subspace(A, 200)
torch.cuda.synchronize()
print(info.tolist())
