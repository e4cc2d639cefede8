"""Filtered evaluation of the real WN18RR test split on a briefly trained model: per-batch time, candidate-list fill and
overflow flag of the tensor-core ranking (workspace words 0 / 1 of rt_score_rank_fused)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rtucker_b200 import ops
from rtucker_b200.engine import SparseTargets
from rtucker_b200.evaluation import rank_batch
from rtucker_b200.train import extract_tensor
dev = torch.device("cuda")
w, graph, label = bench.load_workload("wn18rr")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
run = bench.Runner(w, graph, dev, 1, 0, None, 3, False, steps)
for i in range(steps):
    run.one_step(*[t.to(dev) for t in run.host_batches[i]])
keep = []
orig = ops._ws
ops._ws = lambda n, d: (keep.append(orig(n, d)) or keep[-1])
point = extract_tensor(run.model)
ds = w.get("eval")
for i in range(4):
    f3, boff, bidx = ds.host_batch(np.arange(i * 512, (i + 1) * 512))
    f3, boff, bidx = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (f3, boff, bidx))
    flt = SparseTargets(boff, bidx)
    rank_batch(point, f3, flt); torch.cuda.synchronize()
    del keep[:]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ranks, bce, _ = rank_batch(point, f3, flt); e1.record(); torch.cuda.synchronize()
    big = max(keep, key=lambda t: t.numel())
    scal = big[:16].view(torch.int32).tolist()
    print(f"batch {i}: {e0.elapsed_time(e1):.3f} ms, candidates {scal[0]} ({scal[0] / (512 * 40943) * 100:.2f} % of the pairs), overflow {scal[1]}, "
          f"filter entries {int(boff[-1])}, mean rank {float(ranks.float().mean()):.0f}")
