"""Why does bench.py time the fused kernel slower than tools/v3_graphtime.py?  Same kernel, inputs swapped one at a time."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from rtucker_b200 import ops
from rtucker_b200._lib import lib
dev = torch.device('cuda'); g = torch.Generator().manual_seed(3)
w = bench.WORKLOADS['wn18rr']; B, N, r2 = 512, w['N'], 200
graph = bench.synth_graph(w); feats, off, idx, cnt = graph
items = np.random.default_rng(7).permutation(len(cnt))[:B]
f, boff, bidx = bench.batch_arrays(feats, off, idx, cnt, items)
od, xd = torch.from_numpy(boff).to(dev), torch.from_numpy(bidx).to(dev)
print('batch: nnz', int(boff[-1]), 'max list', int(np.diff(boff).max()))
model = bench.init_params(w); O_model = model.O.weight.data.contiguous().to(dev)
O_tool = torch.linalg.qr(torch.randn(N, r2, generator=g))[0].contiguous().to(dev)
q = (torch.randn(B, r2, generator=g) * 4 * (N / r2) ** 0.5).to(dev)
off2 = torch.arange(0, (B + 1) * 2, 2).int().to(dev); idx2 = torch.randint(0, N, (B * 2,), generator=g).int().to(dev)
ws = torch.empty(int(lib().rt_score_bce_ws_bytes(B, N, r2, 2)) + 16, dtype=torch.uint8, device=dev)
outs = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty(B, r2, device=dev), torch.empty(N, r2, device=dev))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(O, o, x, name):
    run = lambda ph: ops.score_bce_fwd_bwd(q, None, O, o, x, 0.1, variant=2, out=outs, ws=ws, o_absmax=1.0, phases=ph)
    run(7); run(2); run(2); torch.cuda.synchronize()
    e0.record()
    for _ in range(30): run(2)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/30*1e3:.1f} us per launch; |O|max {float(O.abs().max()):.3f}")
t(O_tool, off2, idx2, 'tool O, tool targets ')
t(O_model, off2, idx2, 'model O, tool targets')
t(O_tool, od, xd, 'tool O, bench targets')
t(O_model, od, xd, 'model O, bench targets')
