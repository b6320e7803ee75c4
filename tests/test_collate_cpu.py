"""CPU: the batched scene collate (diffuscene_b200/collate.py) against a per-scene numpy restatement of the reference's
`Diffusion.__getitem__` (threed_front_dataset.py:888-925) and the permutation augmentation (:576-584)."""
import numpy as np
import pytest
import torch

from diffuscene_b200.collate import collate_scenes


def _scene(rng, L, C=23):
    cls = np.eye(C + 1, dtype=np.float32)[rng.integers(0, C - 1, L)]        # real classes; start = col C-1, end = col C
    return {"class_labels": cls, "translations": rng.normal(size=(L, 3)).astype(np.float32),
            "sizes": rng.uniform(0.1, 1, (L, 3)).astype(np.float32), "angles": rng.normal(size=(L, 2)).astype(np.float32),
            "objfeats_32": rng.normal(size=(L, 32)).astype(np.float32)}


def _reference_one(s, max_length):
    out = {}
    cl = np.concatenate([s["class_labels"][:, :-2], s["class_labels"][:, -1:]], axis=-1)
    L, C = cl.shape
    end = np.eye(C)[-1]
    out["class_labels"] = np.vstack([cl, np.tile(end[None], [max_length - L, 1])]).astype(np.float32) * 2.0 - 1.0
    for k in ("translations", "sizes", "angles", "objfeats_32"):
        v = s[k]
        out[k] = np.vstack([v, np.zeros((max_length - v.shape[0], v.shape[1]))]).astype(np.float32)
    return out


def test_padding_and_label_encoding_match_the_reference_per_scene():
    rng = np.random.default_rng(1)
    scenes = [_scene(rng, L) for L in (3, 12, 1, 7, 12)]
    batch = collate_scenes(scenes, max_length=12)
    assert batch["length"].tolist() == [3, 12, 1, 7, 12]
    for b, s in enumerate(scenes):
        ref = _reference_one(s, 12)
        for k, v in ref.items():
            np.testing.assert_array_equal(batch[k][b].numpy(), v)
    with pytest.raises(ValueError):
        collate_scenes(scenes, max_length=8)


def test_permutation_shuffles_objects_inside_each_scene_only():
    rng = np.random.default_rng(2)
    scenes = [_scene(rng, L) for L in (5, 12, 2, 9)]
    g = torch.Generator().manual_seed(0)
    plain = collate_scenes(scenes, max_length=12)
    perm = collate_scenes(scenes, max_length=12, permute=True, generator=g)
    moved = 0
    for b, s in enumerate(scenes):
        L = s["translations"].shape[0]
        # the same set of (translation, size, class) rows, consistently reordered across attributes; padding untouched
        a = torch.cat([plain[k][b, :L] for k in ("translations", "sizes", "class_labels", "objfeats_32")], dim=1)
        p = torch.cat([perm[k][b, :L] for k in ("translations", "sizes", "class_labels", "objfeats_32")], dim=1)
        assert sorted(map(tuple, a.tolist())) == sorted(map(tuple, p.tolist()))
        assert torch.equal(plain["class_labels"][b, L:], perm["class_labels"][b, L:])
        assert torch.equal(perm["translations"][b, L:], torch.zeros(12 - L, 3))
        moved += int(not torch.equal(a, p))
    assert moved >= 2                                   # at least the larger scenes are really shuffled
