"""One fit()/step() from the same state with two score variants: where do they part?  (real WN18RR batch)"""
import os, sys, torch, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rtucker_b200 import asymmetric
from rtucker_b200.data import DeviceEpoch, datasets_from_ids, wn18rr_fixture
from rtucker_b200.optim import FusedLoss
from rtucker_b200.train import extract_tensor
dev = torch.device("cuda:0")
ids = wn18rr_fixture(); train_ds, _, _ = datasets_from_ids(ids, label_smoothing=0.1)
rank = (10, 200, 200)
def make(variant):
    torch.manual_seed(322)
    m = asymmetric.R_TuckER((ids["n_entities"], ids["n_relations"]), rank); m.init(None); m.to(dev)
    o = asymmetric.RSGDwithMomentum([m.core, m.S.weight, m.R.weight, m.O.weight], rank, 109.09, 0.8, score_variant=variant)
    return m, o
va, vb = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 3
(ma, oa), (mb, ob) = make(va), make(vb)
loader = DeviceEpoch(train_ds, 512, dev, shuffle=False, drop_last=True)
rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
for it, (feat, tg) in enumerate(loader):
    if it >= int(sys.argv[3]) if len(sys.argv) > 3 else it >= 6: break
    na = oa.fit(FusedLoss(ma(feat[:, 0], feat[:, 1]), tg, 0.1, 1e-11), extract_tensor(ma))
    nb = ob.fit(FusedLoss(mb(feat[:, 0], feat[:, 1]), tg, 0.1, 1e-11), extract_tensor(mb))
    ea, eb = oa._engine, ob._engine
    print(f"step {it}: |g| {float(na):.6e} vs {float(nb):.6e} (rel {abs(float(na)-float(nb))/float(na):.2e}); "
          f"dS_dir {rel(eb.dS_dir, ea.dS_dir):.2e} dV_S {rel(eb.dV_new[1], ea.dV_new[1]):.2e} dV_O {rel(eb.dV_new[2], ea.dV_new[2]):.2e} "
          f"dV_R {rel(eb.dV_new[0], ea.dV_new[0]):.2e} loss {float(oa.loss):.8f} {float(ob.loss):.8f}")
    oa.step(); ob.step()
    print(f"         after step: core {rel(mb.core.data, ma.core.data):.2e} S {rel(mb.S.weight.data, ma.S.weight.data):.2e} "
          f"O {rel(mb.O.weight.data, ma.O.weight.data):.2e} R {rel(mb.R.weight.data, ma.R.weight.data):.2e}")
