"""CPU: the retrieval oracle (oracle/retrieval_ref.py) against a literal transcription of the reference's
dict / sorted / np.lexsort code (threed_future_dataset.py:28-77) on small catalogues with ties and duplicates."""
import numpy as np

from oracle import retrieval_ref as R


class _Obj:                                           # the three attributes the reference methods touch
    def __init__(self, i, label, size, lat):
        self.i, self.label, self.size, self._lat = i, label, size, lat

    def raw_model_norm_pc_lat32(self):
        return self._lat


def _literal(objects, query_label, query_objfeat, query_size, mode):
    objs = [oi for oi in objects if oi.label == query_label]                       # :25-26
    if not objs:
        return -1
    if mode == 2:                                                                  # :28-35
        mses = {}
        for i, oi in enumerate(objs):
            mses[oi] = np.sum((oi.size - query_size) ** 2, axis=-1)
        return [k for k, v in sorted(mses.items(), key=lambda x: x[1])][0].i
    if mode == 1:                                                                  # :49-59
        mses = {}
        for i, oi in enumerate(objs):
            mses[oi] = np.sum((oi.raw_model_norm_pc_lat32() - query_objfeat) ** 2, axis=-1)
        return [k for k, v in sorted(mses.items(), key=lambda x: x[1])][0].i
    mses_feat, mses_size, keep = [], [], []                                        # :61-77
    for i, oi in enumerate(objs):
        mses_feat.append(np.sum((oi.raw_model_norm_pc_lat32() - query_objfeat) ** 2, axis=-1))
        mses_size.append(np.sum((oi.size - query_size) ** 2, axis=-1))
        keep.append(oi)
    return keep[np.lexsort((mses_feat, mses_size))[0]].i


def make_catalog(rng, M=300, n_classes=7, F=32):
    labels = rng.integers(0, n_classes - 1, M)                   # the last class stays empty
    feats = rng.normal(size=(M, F)).astype(np.float32)
    sizes = rng.uniform(0.1, 2.0, (M, 3)).astype(np.float32)
    sizes[10:40] = sizes[10]                                      # identical sizes: the feature key decides
    feats[20:30] = feats[20]                                      # ... and fully identical entries: the index decides
    labels[10:40] = labels[10]
    return labels, feats, sizes


def test_oracle_matches_literal_reference_code():
    rng = np.random.default_rng(0)
    labels, feats, sizes = make_catalog(rng)
    objects = [_Obj(i, labels[i], sizes[i], feats[i]) for i in range(len(labels))]
    ql = rng.integers(0, 7, 64)
    qf = rng.normal(size=(64, 32)).astype(np.float32)
    qs = rng.uniform(0.1, 2.0, (64, 3)).astype(np.float32)
    ql[:8] = labels[10]
    qs[:8] = sizes[10]                                            # exact size match with 30 candidates
    qf[:4] = feats[20]                                            # ... and exact feature match with 10 of them
    for mode in (0, 1, 2):
        got = R.retrieve_batch(labels, feats, sizes, ql, qf, qs, mode)
        want = np.asarray([_literal(objects, ql[q], qf[q], qs[q], mode) for q in range(64)])
        assert np.array_equal(got, want), mode
    assert (R.retrieve_batch(labels, feats, sizes, ql, qf, qs, 0)[ql == 6] == -1).all()
    assert R.retrieve_batch(labels, feats, sizes, ql, qf, qs, 0)[0] == 20        # first of the identical entries
