"""Role-level cycle counters of apply_tc_kernel (build: make -C r-tucker_b200/csrc prof PROF_SRC=apply_tc PROF_DEF=RT_APPLY_PROF;
run: [RT_APPLY_DEBUG=bits] python tools/with_lib.py tools/_prof/librt_prof.so tools/apply_prof.py).
RT_APPLY_DEBUG bits: 1 no X loads, 2 no K-image loads after the first ring fill, 4 one MMA per block, 8 ld.global.cg for X."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rtucker_b200 import ops
from microbench import timeit
dev = torch.device('cuda'); N, r = 40943, 200
torch.manual_seed(0)
U, V, W = (torch.randn(N, r, device=dev) for _ in range(3))
K = torch.randn(r, r, device=dev, dtype=torch.float64)
Y = torch.empty(N, r, device=dev); Y2 = torch.empty(N, r, device=dev)
fn = lambda: ops.apply_multi([(Y, None, None, [(U, K), (V, K), (W, K)]), (Y2, None, None, [(U, K), (V, K), (W, K)])])
fn(); torch.cuda.synchronize()
print("---- debug", os.environ.get("RT_APPLY_DEBUG", "0"))
