"""Sparse-target data pipeline.

Replaces ``KG_dataset.__getitem__`` + default collate (reference src/data/Dataset.py:42-53,
train.py:226-236), which allocate a dense ``zeros(n_ent)`` row per sample (84 MB per WN18RR batch):
the (s, r) -> objects vocabulary is built once on the host as CSR, a batch is (features, CSR slice),
and only a few KB cross PCIe per step.  Item order follows the reference: train items are the unique
(s, r) pairs in first-occurrence order (Dataset.py:14-15,29-34), test items are the triples.
"""
from collections import OrderedDict

import numpy as np
import torch

from .engine import SparseTargets


class SparseKGDataset:
    def __init__(self, triples, n_entities, all_triples=None, test_set=False, label_smoothing=0.0):
        """triples: int array [T,3] of (s, r, o) ids for this split (reverse triples already added,
        as Data.load_data does); all_triples: every split (the filter set) when test_set."""
        triples = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        self.n_entities = int(n_entities)
        self.test_set = bool(test_set)
        self.label_smoothing = 0.0 if test_set else float(label_smoothing or 0.0)
        vocab_src = np.asarray(all_triples, dtype=np.int64).reshape(-1, 3) if test_set else triples
        vocab = OrderedDict()
        for s, r, o in vocab_src.tolist():
            vocab.setdefault((s, r), []).append(o)
        if test_set:
            self.features = triples.astype(np.int32)                      # (s, r, o) per item
            keys = [(s, r) for s, r, _ in triples.tolist()]
        else:
            pairs = OrderedDict()
            for s, r, _ in triples.tolist():
                pairs.setdefault((s, r), None)
            keys = list(pairs.keys())
            self.features = np.asarray(keys, dtype=np.int32).reshape(-1, 2)
        lists = [np.unique(np.asarray(vocab[k], dtype=np.int32)) for k in keys]  # unique + ascending
        self.counts = np.asarray([len(x) for x in lists], dtype=np.int64)
        self.off = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum(self.counts, out=self.off[1:])
        self.idx = np.concatenate(lists).astype(np.int32) if lists else np.zeros(0, np.int32)

    def __len__(self):
        return self.features.shape[0]

    def num_triples(self, items=None):
        """Number of (s,r,o) targets covered by ``items`` (all items if None): the 'train triples'."""
        return int(self.counts.sum() if items is None else self.counts[items].sum())

    def host_batch(self, items):
        items = np.asarray(items, dtype=np.int64)
        cnt = self.counts[items]
        off = np.zeros(len(items) + 1, dtype=np.int32)
        np.cumsum(cnt, out=off[1:])
        starts = self.off[items]
        gather = np.repeat(starts - off[:-1], cnt) + np.arange(int(off[-1]), dtype=np.int64)
        return self.features[items], off, self.idx[gather]

    def batch(self, items, device, pinned=None):
        """(features int32 [B, 2|3], SparseTargets) on ``device``; host->device copies are async from
        pinned staging buffers.  Returns also the number of bytes copied."""
        feat, off, idx = self.host_batch(items)
        tensors = []
        nbytes = 0
        for a in (feat, off, idx):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if device.type == "cuda":
                t = t.pin_memory().to(device, non_blocking=True)
            nbytes += a.nbytes
            tensors.append(t)
        return tensors[0], SparseTargets(tensors[1], tensors[2]), nbytes

    def batches(self, batch_size, device, shuffle=False, drop_last=False, generator=None):
        n = len(self)
        order = torch.randperm(n, generator=generator).numpy() if shuffle else np.arange(n)
        end = n - (n % batch_size) if drop_last else n
        for lo in range(0, end, batch_size):
            items = order[lo:min(lo + batch_size, end)]
            yield (*self.batch(items, device), items)


def from_reference_data(data, split, test_set=False, label_smoothing=0.0):
    """Build a SparseKGDataset from an object with the reference ``Data`` fields
    (src/data/Data.py: entities, relations, train_data/valid_data/test_data, data)."""
    ent = {e: i for i, e in enumerate(data.entities)}
    rel = {r: i for i, r in enumerate(data.relations)}

    def ids(rows):
        return np.asarray([(ent[a], rel[b], ent[c]) for a, b, c in rows], dtype=np.int64).reshape(-1, 3)
    rows = getattr(data, f"{split}_data")
    return SparseKGDataset(ids(rows), len(data.entities), all_triples=ids(data.data) if test_set else None,
                           test_set=test_set, label_smoothing=label_smoothing)
