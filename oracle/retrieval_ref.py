"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's
catalogue retrieval, scene_synthesis/datasets/threed_future_dataset.py:28-77.

`closest_*` follow the reference functions line by line on plain numpy arrays instead of `ThreedFutureModel`
objects (an "object" here is its index into the catalogue arrays): filter by label in catalogue order (:25-26),
float32 `np.sum((a - b) ** 2, axis=-1)` per candidate, then
  * _to_box (:28-35) / _to_objfeats (:49-59): `sorted(mses.items(), key=value)` -> first  (stable: first minimum)
  * _to_objfeats_and_size (:61-77): `np.lexsort((mses_feat, mses_size))` -> first  (primary key: size mse)
Pinned by tests/test_retrieval_cpu.py against a literal dict / sorted / lexsort transcription on small catalogues.
"""
from __future__ import annotations

import numpy as np


def _filter(cat_labels: np.ndarray, label: int) -> np.ndarray:
    return np.nonzero(cat_labels == label)[0]                     # catalogue order, like the list comprehension (:25-26)


def closest_to_box(cat_labels, cat_sizes, label, size) -> int:
    idx = _filter(cat_labels, label)
    if idx.size == 0:
        return -1
    mses = [np.sum((cat_sizes[i] - size) ** 2, axis=-1) for i in idx]
    order = sorted(range(len(idx)), key=lambda j: mses[j])        # stable, like sorted(dict.items())
    return int(idx[order[0]])


def closest_to_objfeats(cat_labels, cat_feats, label, feat) -> int:
    idx = _filter(cat_labels, label)
    if idx.size == 0:
        return -1
    mses = [np.sum((cat_feats[i] - feat) ** 2, axis=-1) for i in idx]
    order = sorted(range(len(idx)), key=lambda j: mses[j])
    return int(idx[order[0]])


def closest_to_objfeats_and_size(cat_labels, cat_feats, cat_sizes, label, feat, size) -> int:
    idx = _filter(cat_labels, label)
    if idx.size == 0:
        return -1
    mses_feat = [np.sum((cat_feats[i] - feat) ** 2, axis=-1) for i in idx]
    mses_size = [np.sum((cat_sizes[i] - size) ** 2, axis=-1) for i in idx]
    ind = np.lexsort((mses_feat, mses_size))
    return int(idx[ind[0]])


def retrieve_batch(cat_labels, cat_feats, cat_sizes, q_labels, q_feats, q_sizes, mode: int) -> np.ndarray:
    """All queries, one at a time like the reference's generation script does per object."""
    out = np.empty(len(q_labels), dtype=np.int64)
    for q in range(len(q_labels)):
        if mode == 0:
            out[q] = closest_to_objfeats_and_size(cat_labels, cat_feats, cat_sizes, q_labels[q], q_feats[q], q_sizes[q])
        elif mode == 1:
            out[q] = closest_to_objfeats(cat_labels, cat_feats, q_labels[q], q_feats[q])
        else:
            out[q] = closest_to_box(cat_labels, cat_sizes, q_labels[q], q_sizes[q])
    return out
