import sys, torch
sys.path.insert(0,'/root/repo')
from rtucker_b200 import ops
n=400; torch.manual_seed(0)
X=torch.randn(n,4*n,dtype=torch.float64,device='cuda'); A=X@X.T
w,V=ops.eigh(A); torch.cuda.synchronize(); print(float(w[0]))
