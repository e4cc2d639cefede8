import sys, torch
sys.path.insert(0, '/root/repo')
from rtucker_b200._lib import lib, ptr, stream_ptr, check
out = torch.zeros(2, dtype=torch.int64, device='cuda')
for N in (96, 128, 208, 256):
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 1), (1, 0)):
        for reps in (64, 512):
            check(lib().rt_mma_probe(N, 8, a_mn, b_mn, reps, ptr(out), stream_ptr()), 'probe'); torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"N={N:3d} a_mn={a_mn} b_mn={b_mn} reps={reps:3d}: issue {o[0]/reps:6.1f} cyc/MMA, complete {o[1]/reps:6.1f} cyc/MMA (floor {128*N/256:.0f})")
