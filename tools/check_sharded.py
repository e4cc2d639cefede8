"""torchrun --nproc-per-node 2 tools/check_sharded.py : entity-sharded engine over NCCL vs the unsharded engine
(same seeds) on real GPUs; rank 0 prints the deviations."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
from rtucker_b200.engine import SparseTargets, StepEngine
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
N, M, rk, B = 6000, 14, (8, 48, 48), 256
g = torch.Generator().manual_seed(3)
q = lambda a, b: torch.linalg.qr(torch.randn(a, b, generator=g))[0].contiguous()
core = 900 * torch.randn(rk, generator=g); R, S, O = q(M, rk[0]), q(N, rk[1]), q(N, rk[2])
per = N // world; lo, hi = rank * per, (rank + 1) * per
P = torch.nn.Parameter
def mk(sl, group):
    pc = P(core.clone().to(dev)); pf = [P(R.clone().to(dev)), P(S[sl].clone().to(dev)), P(O[sl].clone().to(dev))]
    return pc, pf, StepEngine(pc, pf, False, B, 0.8, group=group, n_total=N, n_begin=sl.start or 0)
pc_s, pf_s, eng_s = mk(slice(lo, hi), dist.group.WORLD)
pc_u, pf_u, eng_u = mk(slice(0, N), None)
out = []
for it in range(3):
    sub = torch.randint(0, N, (B,), generator=g).int().to(dev); rel = torch.randint(0, M, (B,), generator=g).int().to(dev)
    off = torch.arange(0, 2 * B + 1, 2).int().to(dev); idx = torch.randint(0, N, (2 * B,), generator=g).int().to(dev)
    t = SparseTargets(off, idx)
    ns = eng_s.fit(rel, sub, t, 0.1, 1e-9); eng_s.step(400.0)
    nu = eng_u.fit(rel, sub, t, 0.1, 1e-9); eng_u.step(400.0)
    out.append((float(eng_s.loss), float(eng_u.loss), float(ns), float(nu)))
torch.cuda.synchronize()
dc = float((pc_s.data - pc_u.data).norm() / pc_u.data.norm())
do = float((pf_s[2].data - pf_u[2].data[lo:hi]).norm() / pf_u[2].data[lo:hi].norm())
if rank == 0:
    for a in out: print('loss sharded %.9f unsharded %.9f | norm %.6e %.6e' % a)
    print('core rel diff %.2e, O-shard rel diff %.2e' % (dc, do))
dist.destroy_process_group()
