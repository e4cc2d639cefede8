"""Kernel-only time of score_v3_kernel measured three ways: eager back-to-back launches, a CUDA graph of 20 launches,
and single launches bracketed by events."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from rtucker_b200 import ops
from rtucker_b200._lib import lib
dev = torch.device('cuda'); g = torch.Generator().manual_seed(3)
B, N, r2 = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 40943, 200
O = torch.linalg.qr(torch.randn(N, r2, generator=g))[0].contiguous().to(dev)
q = (torch.randn(B, r2, generator=g) * 4 * (N / r2) ** 0.5).to(dev)
off = torch.arange(0, (B + 1) * 2, 2).int().to(dev); idx = torch.randint(0, N, (B * 2,), generator=g).int().to(dev)
ws = torch.empty(int(lib().rt_score_bce_ws_bytes(B, N, r2, 2)) + 16, dtype=torch.uint8, device=dev)
outs = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty(B, r2, device=dev), torch.empty(N, r2, device=dev))
run = lambda ph: ops.score_bce_fwd_bwd(q, None, O, off, idx, 0.1, variant=2, out=outs, ws=ws, o_absmax=1.0, phases=ph)
run(7); run(2); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(50): run(2)
e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"eager back-to-back: {e0.elapsed_time(e1)/50*1e3:.1f} us per launch (host enqueue {1e6*(t1-t0)/50:.1f} us per call)")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(2); torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(20): run(2)
    gph.replay(); torch.cuda.synchronize()
    e0.record(); gph.replay(); gph.replay(); e1.record(); torch.cuda.synchronize()
print(f"graph of 20 launches: {e0.elapsed_time(e1)/40*1e3:.1f} us per launch")
