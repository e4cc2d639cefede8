#!/usr/bin/env python
"""Generates tests/golden/wn18rr_trajectory.npz: loss and ||rgrad|| of the first STEPS training steps on the REAL
WN18RR batches (tests/golden/wn18rr_ids.npz, dataset order, batch 512) as the reference executes them --
oracle/reference_step.py, the port of train.py:79 + src/model/asymmetric/optim.py:74-114 (autodiff through the
rank-2r construct, QR + SVD rounding) on the restated toolbox -- in fp64, from R_TuckER.init at seed 322 (README
recipe: rank (10, 200, 200), momentum 0.8, label smoothing 0.1; lr 109.09 = OneCycleLR's first epoch at HEAD,
reg 1e-11 = configs/base_config.py).  Run in the build container (minutes of CPU):
    python tests/golden/make_trajectory.py [steps]
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reference_step as RS  # noqa: E402
from rtucker_b200 import asymmetric  # noqa: E402
from rtucker_b200.data import datasets_from_ids, wn18rr_fixture  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 50
SEED, RANK, B, LS, LR, REG, BETA = 322, (10, 200, 200), 512, 0.1, 600 / 5.5, 1e-11, 0.8


def main():
    ids = wn18rr_fixture()
    train, _, _ = datasets_from_ids(ids, label_smoothing=LS)
    np.random.seed(SEED)
    torch.manual_seed(SEED)
    model = asymmetric.R_TuckER((ids["n_entities"], ids["n_relations"]), RANK)
    model.init(None)
    torch.set_default_dtype(torch.float64)
    st = RS.ReferenceStepper(model.core.data.double(), model.R.weight.data.double(), model.S.weight.data.double(),
                             model.O.weight.data.double(), BETA)
    loss, norm = [], []
    for k in range(STEPS):
        t0 = time.time()
        feat, off, idx = train.host_batch(np.arange(k * B, (k + 1) * B))
        f = torch.from_numpy(feat).long()
        tg = RS.dense_targets(ids["n_entities"], torch.from_numpy(off).long(), torch.from_numpy(idx).long(), LS).double()
        n = st.train_step(f[:, 0], f[:, 1], tg, REG, LR)
        loss.append(float(st.loss))
        norm.append(float(n))
        print(k, loss[-1], norm[-1], f"{time.time() - t0:.1f}s", flush=True)
    np.savez_compressed(os.path.join(HERE, "wn18rr_trajectory.npz"), loss=np.asarray(loss), norm=np.asarray(norm),
                        seed=SEED, rank=np.asarray(RANK), batch=B, ls=LS, lr=LR, reg=REG, beta=BETA, steps=STEPS)


if __name__ == "__main__":
    main()
