import sys, torch, ctypes as C
sys.path.insert(0,'/root/repo')
from rtucker_b200 import ops
from rtucker_b200._lib import lib
L=lib(); L.rt_eigh_visit_profile.argtypes=[C.POINTER(C.c_longlong)]; L.rt_eigh_visit_profile.restype=C.c_int
buf=(C.c_longlong*8)()
n=400; X=torch.randn(n,4*n,dtype=torch.float64,device='cuda'); A=X@X.T
L.rt_eigh_visit_profile(buf)
ops.eigh(A); torch.cuda.synchronize(); L.rt_eigh_visit_profile(buf)
tn,ta,rounds,tw=buf[0],buf[1],buf[2],buf[3]
print(f'inner rounds {rounds}: next_pairs {tn/rounds:.0f} clk/round, apply {ta/rounds:.0f} clk/round, barrier wait (apply thread) {tw/rounds:.0f} clk/round')
