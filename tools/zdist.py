"""Logit distribution seen by the fused score kernel after a few bench-like training steps."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from rtucker_b200 import asymmetric, ops
from rtucker_b200.engine import SparseTargets
from rtucker_b200.optim import FusedLoss
w = bench.WORKLOADS['wn18rr']; dev = torch.device('cuda'); N = w['N']
graph = bench.synth_graph(w); feats, off, idx, cnt = graph
model = bench.init_params(w); model.to(dev)
opt = asymmetric.RSGDwithMomentum([model.core, model.S.weight, model.R.weight, model.O.weight], w['rank'], bench.LR, 0.8, score_variant=2)
opt.param_groups[0]['lr'] = bench.LR
order = np.random.default_rng(7).permutation(len(cnt))
def stats(tag):
    items = order[:512]
    f, boff, bidx = bench.batch_arrays(feats, off, idx, cnt, items)
    fd = torch.from_numpy(f).to(dev)
    r_rows = ops.gather_rows(model.R.weight.data, fd[:, 1].contiguous()); s_rows = ops.gather_rows(model.S.weight.data, fd[:, 0].contiguous())
    q = ops.query_fwd(model.core.data, r_rows, s_rows)
    O = model.O.weight.data
    Z = q @ O.T
    qq = torch.randn(512, 200, device=dev) * 4.0 * (N / 200) ** 0.5
    Z2 = qq @ O.T
    for name, z in (('model q', Z), ('synthetic q', Z2)):
        out = ((z > 16.6) | (z < -27.7)).float()
        grp = out[:, :N // 16 * 16].view(512, -1, 16).amax(2)              # per thread group
        wgrp = grp.view(16, 32, -1).amax(1)                                 # per warp (32 rows)
        print(f"{tag} {name}: z mean {float(z.mean()):.2f} std {float(z.std()):.2f} min {float(z.min()):.1f} max {float(z.max()):.1f}; out-of-window elems {float(out.mean()):.4f}, warp-groups {float(wgrp.mean()):.4f}; |q| {float(q.norm(dim=1).mean()):.1f} core norm {float(model.core.data.norm()):.2f}")
    rn = O.norm(dim=1); print(f"   O row norms: mean {float(rn.mean()):.4f} max {float(rn.max()):.4f}")
stats('init')
for it in range(15):
    items = order[it * 512:(it + 1) * 512]
    f, boff, bidx = bench.batch_arrays(feats, off, idx, cnt, items)
    fd, od, xd = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (f, boff, bidx))
    opt.fit(FusedLoss(model(fd[:, 0], fd[:, 1]), SparseTargets(od, xd), 0.1, bench.REG), None); opt.step()
torch.cuda.synchronize()
stats('after 15 steps')
