"""In-kernel phase profile of score_v3_kernel (clock64 per role).  python tools/v3_prof.py [N]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from rtucker_b200 import ops
from rtucker_b200._lib import lib, ptr
dev = torch.device('cuda'); g = torch.Generator().manual_seed(3)
B, N, r2 = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 40943, 200
O = torch.linalg.qr(torch.randn(N, r2, generator=g))[0].contiguous().to(dev)
q = (torch.randn(B, r2, generator=g) * 4 * (N / r2) ** 0.5).to(dev)
off = torch.arange(0, (B + 1) * 2, 2).int().to(dev); idx = torch.randint(0, N, (B * 2,), generator=g).int().to(dev)
import os
def timeit(fn, n=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
ws = torch.empty(int(lib().rt_score_bce_ws_bytes(B, N, r2, 2)) + 16, dtype=torch.uint8, device=dev)
outs = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty(B, r2, device=dev), torch.empty(N, r2, device=dev))
print("flush mode", os.environ.get("FLUSH", "0"), "time us:", 1e3 * timeit(lambda: ops.score_bce_fwd_bwd(q, None, O, off, idx, 0.1, variant=2, out=outs, ws=ws, o_absmax=1.0)))
ref = ops.score_bce_fwd_bwd(q, q, O, off, idx, 0.1, variant=0)
print("rel err H", float((outs[1] - ref[1]).norm() / ref[1].norm()), "dO", float((outs[2] - ref[2]).norm() / ref[2].norm()))
for _ in range(2):
    ops.score_bce_fwd_bwd(q, None, O, off, idx, 0.1, variant=2, o_absmax=1.0)
prof = torch.zeros(148, 4, 10, dtype=torch.int64, device=dev)
lib().rt_score_v3_set_profile(ptr(prof))
ops.score_bce_fwd_bwd(q, None, O, off, idx, 0.1, variant=2, o_absmax=1.0)
torch.cuda.synchronize(); lib().rt_score_v3_set_profile(None)
p = prof.double().cpu()
names = [["wait oempty", "wait qempty", "issue/other"] + [""] * 7,
         ["wait ofull", "wait qfull", "wait zfree", "issue G1", "wait gfull", "wait d2free", "issue G2", "wait d3free", "issue G3", ""],
         ["mask", "wait zfull", "tmem ld", "math", "wait gfree", "store G", "", "", "", ""],
         ["wait d2full", "D2 red.add", "D2 tmem ld wait", "D2 first st", "", "", "", "wait d3full", "d3 stores", "d3 tmem ld wait"]]
for r, role in enumerate(["producer", "mma", "epilogue(w2)", "flush leader"]):
    tot = p[:, r].sum(1).mean()
    print(f"{role}: total {tot:.0f} cycles")
    for i, n in enumerate(names[r]):
        if n: print(f"   {n:14s} mean {p[:, r, i].mean():10.0f}  max {p[:, r, i].max():10.0f}  ({100 * p[:, r, i].mean() / tot:.1f}%)")
