import sys, torch
sys.path.insert(0,'/root/repo')
from rtucker_b200._lib import lib, ptr, stream_ptr, check
N, K, a_mn, b_mn, flags = (int(x) for x in sys.argv[1:6])
dev=torch.device('cuda'); torch.manual_seed(0)
A=torch.randn(128,K,device=dev); B=torch.randn(N,K,device=dev)
Ain = A.t().contiguous() if a_mn else A
Bin = B.t().contiguous() if b_mn else B
D=torch.zeros(128,N,device=dev)
check(lib().rt_tc_selftest(ptr(Ain),ptr(Bin),ptr(D),N,K,a_mn,b_mn,flags,stream_ptr()),'selftest')
torch.cuda.synchronize()
ref=A.double()@B.double().t()
err=float((D.double()-ref).norm()/ref.norm())
print(f"N={N} K={K} a_mn={a_mn} b_mn={b_mn} flags={flags} relerr={err:.3e}", "OK" if err<3e-3 else "WRONG")
