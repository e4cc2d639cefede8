"""Host-side logic on the CPU: the step engine's state machine (driven through the CPU stand-in for the
C-ABI ops, tests/cpu_ops.py) against the reference golden trajectories; model init parity; the sparse-target
data pipeline against the reference KG_dataset."""
import numpy as np
import pytest
import torch

import cpu_ops
import golden_util
import ref_harness
from rtucker_b200.engine import SparseTargets, StepEngine

f64 = torch.float64


@pytest.mark.parametrize("name,beta", [("steps_asym_rsgd.npz", 0.8), ("steps_sym_rsgd.npz", 0.8), ("steps_sym_rgd.npz", None)])
def test_engine_state_machine_matches_reference_golden(name, beta):
    fx = golden_util.step_fixture(name)
    i = fx["init"]
    P = torch.nn.Parameter
    core = P(i["core"].clone())
    fs = [P(i["R"].clone()), P(i["E"].clone())] if fx["sym"] else [P(i["R"].clone()), P(i["S"].clone()), P(i["O"].clone())]
    eng = StepEngine(core, fs, fx["sym"], 24, beta, ops=cpu_ops)
    for k, (rel, sub, off, idx) in enumerate(fx["batches"]):
        nrm = eng.fit(rel.int(), sub.int(), SparseTargets(off.int(), idx.int()), fx["ls"], fx["reg"])
        eng.step(fx["lr"])
        assert abs(float(eng.loss) - fx["loss"][k]) / fx["loss"][k] < 1e-10
        assert abs(float(nrm) - fx["norm"][k]) / fx["norm"][k] < 1e-9
    import analytic as A
    f = [p.data for p in fs]
    X = A.Point(core.data, [f[0], f[1], f[1] if fx["sym"] else f[2]], fx["sym"]).to_dense()
    assert float((X - fx["X_final"]).norm() / fx["X_final"].norm()) < 1e-9


@pytest.mark.parametrize("kind", ["asym", "sym"])
def test_model_init_matches_reference(kind):
    """Same seed, same RNG call order as R_TuckER.init => identical initial parameters."""
    from rtucker_b200 import asymmetric, symmetric
    z = golden_util.load(f"scores_{kind}.npz")
    np.random.seed(20)
    torch.manual_seed(20)            # set_random_seed(20) of the reference (utils.py:8-12)
    mod = symmetric if kind == "sym" else asymmetric
    model = mod.R_TuckER((int(z["N"]), int(z["M"])), tuple(int(x) for x in z["rank"]))
    model.init(None)
    assert np.array_equal(model.core.data.numpy(), z["core"])
    assert np.allclose(model.R.weight.data.numpy(), z["R"], atol=0)
    if kind == "sym":
        assert np.array_equal(model.E.weight.data.numpy(), z["E"])
        assert list(model.state_dict().keys()) == ["core", "E.weight", "R.weight"]
    else:
        assert np.array_equal(model.S.weight.data.numpy(), z["S"]) and np.array_equal(model.O.weight.data.numpy(), z["O"])
        assert list(model.state_dict().keys()) == ["core", "S.weight", "R.weight", "O.weight"]


def test_sparse_dataset_semantics():
    from rtucker_b200.data import SparseKGDataset
    triples = np.array([[0, 0, 1], [0, 0, 2], [3, 1, 0], [0, 0, 1], [2, 1, 4]])
    allt = np.concatenate([triples, np.array([[0, 0, 4], [3, 1, 1]])])
    tr = SparseKGDataset(triples, 5, label_smoothing=0.1)
    assert len(tr) == 3 and tr.features.tolist() == [[0, 0], [3, 1], [2, 1]]     # unique (s,r), first-seen order
    f, off, idx = tr.host_batch([0, 2])
    assert off.tolist() == [0, 2, 3] and idx.tolist() == [1, 2, 4]                # duplicates removed, ascending
    assert tr.num_triples() == 4
    te = SparseKGDataset(triples[:3], 5, all_triples=allt, test_set=True)
    assert len(te) == 3 and te.features.tolist() == triples[:3].tolist() and te.label_smoothing == 0.0
    f, off, idx = te.host_batch([0, 2])
    assert idx[off[0]:off[1]].tolist() == [1, 2, 4] and idx[off[1]:off[2]].tolist() == [0, 1]  # filter over ALL splits


@pytest.mark.skipif(not ref_harness.available(), reason="reference checkout (with data/WN18RR) not present")
def test_sparse_dataset_matches_reference_kg_dataset_on_wn18rr():
    import os
    from rtucker_b200.data import from_reference_data
    z = golden_util.load("dataset_wn18rr.npz")
    ns = ref_harness.load("asymmetric", "rsgd")
    data = ns.Data(os.path.join(ref_harness.REF_DIR, "data", "WN18RR"), reverse=True)
    tr = from_reference_data(data, "train", label_smoothing=0.1)
    va = from_reference_data(data, "valid", test_set=True)
    assert tr.n_entities == int(z["n_entities"]) == 40943 and len(tr) == int(z["n_train_items"]) == 103509
    assert len(va) == int(z["n_valid_items"]) and tr.num_triples() == 173670
    for name, ds in (("train", tr), ("valid", va)):
        for k, item in enumerate(z[f"{name}_items"]):
            f, off, idx = ds.host_batch([int(item)])
            assert f[0].tolist() == z[f"{name}_features"][k].tolist()
            assert idx.tolist() == sorted(z[f"{name}_targets"][k].tolist())


def test_fit_requires_fused_loss():
    from rtucker_b200 import asymmetric
    torch.manual_seed(0)
    m = asymmetric.R_TuckER((20, 4), (2, 3, 3))
    m.init(None)
    opt = asymmetric.RSGDwithMomentum([m.core, m.S.weight, m.R.weight, m.O.weight], (2, 3, 3), 10.0, 0.8)
    with pytest.raises(TypeError):
        opt.fit(lambda T: T.norm(), None)
    with pytest.raises(RuntimeError):
        opt.step()
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=600, total_steps=500, pct_start=100 / 500, div_factor=5.5,
                                                cycle_momentum=False, anneal_strategy="linear")
    assert abs(opt.param_groups[0]["lr"] - 600 / 5.5) < 1e-9      # train.py:213-215 drives param_groups[0]["lr"]
    assert "lr" in opt.state_dict()["param_groups"][0] and sched is not None
