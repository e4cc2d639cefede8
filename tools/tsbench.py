import sys, torch
sys.path.insert(0,'/root/repo')
from rtucker_b200 import ops
sys.path.insert(0,'/root/repo/tools')
from microbench import timeit
dev=torch.device('cuda'); N,r=40943,200
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
U,V,W=(torch.randn(N,r,device=dev) for _ in range(3)); K=torch.randn(r,r,device=dev,dtype=torch.float64); Y=torch.empty(N,r,device=dev)
for tc in (False, True):
    print('tc',tc,'gram %.3f ms'%timeit(lambda: ops.gram(U,V,tc=tc),flush=flush),
          'apply1 %.3f'%timeit(lambda: ops.apply(Y,Y,None,[(U,K)],tc=tc),flush=flush),
          'apply2 %.3f'%timeit(lambda: ops.apply(Y,None,None,[(U,K),(V,K)],tc=tc),flush=flush),
          'apply3 %.3f'%timeit(lambda: ops.apply(Y,Y,None,[(U,K),(V,K),(W,K)],tc=tc),flush=flush))
ref=(U.double().T@V.double())
for tc in (False,True):
    g=ops.gram(U,V,tc=tc); print('gram relerr tc',tc, float((g-ref).norm()/ref.norm()))
ref=U.double()@K+V.double()@K
for tc in (False,True):
    ops.apply(Y,None,None,[(U,K),(V,K)],tc=tc); print('apply2 relerr tc',tc, float((Y.double()-ref).norm()/ref.norm()))
