"""Checkpoint / resume.

The reference's ``StateDict.save`` writes model, losses, metrics and last_epoch but DROPS the optimiser and
scheduler dictionaries it was handed (src/utils/storage.py:70-78, train.py:154-159), RSGD's momentum is not in
``optimizer.state_dict()`` at all, and ``StateDict.load`` cannot rebuild the dataclass (storage.py:80-83).  Here a
checkpoint is one plain dictionary of tensors and numbers (loadable with ``weights_only=True``):

  model       model.state_dict()  -- the reference's keys: core, S.weight, R.weight, O.weight / core, E.weight, R.weight
  optimizer   optimizer.state_dict() incl. "rtucker_engine": kept direction, its base point, transport Grams,
              Adam moments, hyper-parameters (per rank: the entity rows this rank owns)
  scheduler   scheduler.state_dict()
  last_epoch, history (list of per-epoch records: what Losses / Metrics accumulate, storage.py:8-58)

A resumed run continues the uninterrupted trajectory bit for bit (tests/test_gpu_api.py::test_checkpoint_resume).
"""
import os

import torch


def save(path, model, optimizer=None, scheduler=None, last_epoch=0, history=None):
    state = {
        "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizer": optimizer.state_dict() if optimizer is not None else None,
        "scheduler": scheduler.state_dict() if scheduler is not None else None,
        "last_epoch": int(last_epoch),
        "history": list(history or []),
    }
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)           # a crash never leaves a truncated snapshot behind
    return state


def load(path, model=None, optimizer=None, scheduler=None, map_location="cpu"):
    """Returns the stored dictionary; restores whichever of model / optimizer / scheduler are given."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    if model is not None:
        model.load_state_dict(state["model"])
    if optimizer is not None and state.get("optimizer") is not None:
        optimizer.load_state_dict(state["optimizer"])
    if scheduler is not None and state.get("scheduler") is not None:
        scheduler.load_state_dict(state["scheduler"])
    return state
