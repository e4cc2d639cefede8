import sys, torch
sys.path.insert(0,'/root/repo')
from rtucker_b200._lib import lib, ptr, stream_ptr, check
N, K, a_mn, b_mn, flags = (int(x) for x in sys.argv[1:6])
dev=torch.device('cuda'); torch.manual_seed(0)
# structured inputs: A[m,k] = 1 if k == m % K (selector), so D[m,n] = B[n, m % K]
A=torch.zeros(128,K,device=dev); A[torch.arange(128), torch.arange(128) % K] = 1.0
B=(torch.arange(N,device=dev).float()[:,None]*100 + torch.arange(K,device=dev).float()[None,:])
Ain = A.t().contiguous() if a_mn else A
Bin = B.t().contiguous() if b_mn else B
D=torch.zeros(128,N,device=dev)
check(lib().rt_tc_selftest(ptr(Ain),ptr(Bin),ptr(D),N,K,a_mn,b_mn,flags,stream_ptr()),'selftest')
torch.cuda.synchronize()
ref=A@B.t()
print(f"cfg N={N} K={K} a_mn={a_mn} b_mn={b_mn} flags={flags} maxabs={float(D.abs().max()):.1f} match={bool(torch.allclose(D,ref))}")
print(" D[0:4,0:6]  ", D[0:4,0:6].tolist())
print(" ref[0:4,0:6]", ref[0:4,0:6].tolist())
print(" D[9,0:6]", D[9,0:6].tolist(), " ref[9,0:6]", ref[9,0:6].tolist())
